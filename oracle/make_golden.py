"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container (the only place /root/reference exists):

    python oracle/make_golden.py

Every vector below is produced by the reference's own objects (hypernet_attention.HyperNet, AttentionGru,
hypernet.HyperNet, DecoderGRU) imported through oracle/ref_harness.py -- not by the oracle port.  The port
(oracle/caption_hn_oracle.py) and the CUDA path are then both checked against these files.

Cases
  attn_flickr : Variant B, style vector [1,E] (hypernet_attention.py:139-142), teacher-forced + greedy, literal grads
  attn_cc     : Variant B, cc=True, 1-D one-hot style [he] (cc_train_hypernet.py:141-144), teacher-forced + greedy
  pooled_l1   : Variant A, L=1, DecoderGRU.forward + infer (V must be 9684, later.py:449)
  pooled_l2   : Variant A, L=2 (exercises the offset-0 slicing of utils.py:45,68)
"flow" gradients (hypernet heads receiving gradient) come from the reference with ONE change made here in the
harness, not in the reference: set_all_parameters without the nn.Parameter(...) wrapper of utils.py:57.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _sd(model):
    return {k: v.detach().clone() for k, v in model.state_dict().items()
            if not (k.startswith("image_encoder.") and not k.startswith("image_encoder.fc."))}


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def _flow_set_all_parameters(module, theta):
    """utils.py:44-69 minus the nn.Parameter wrapper at :57 (keeps the autograd graph)."""
    count = 0
    for name in module.registered_parameters_name:
        if name in ("weight", "bias"):
            continue
        a = count
        b = a + getattr(module, name).numel()
        t = torch.reshape(theta[0, a:b], getattr(module, name).shape)
        if name in module._parameters:
            del module._parameters[name]
        object.__setattr__(module, name, t)
        count += t.numel()
    for name in [k for k in module._modules]:
        if name in ("embed", "fc_out"):
            continue
        count += _flow_set_all_parameters(module._modules[name], theta)
    return count


def case_attention(ref, name, cc, style, Fo=16, E=12, H=20, V=50, he=10, B=2, T=6):
    torch.manual_seed(0)
    model = ref.HyperNetAttention(Fo, E, H, V, ref.vocab, cc=cc, hyper_emb=he)
    sd = _sd(model)
    g = torch.Generator().manual_seed(1234)
    features = torch.randn(B, 49, 2048, generator=g)
    caps = torch.randint(4, V, (B, T), generator=g)
    caps[:, 0] = 1
    caps[1, T - 2] = 2
    caps[1, T - 1] = 0
    out = {"features": features, "captions": caps, "style": style}
    out.update({"sd/" + k: v for k, v in sd.items() if not k.startswith("captioner.gru.")})

    # literal mode, teacher-forced (cc_train_hypernet.py:150-153)
    np.random.seed(0)
    captioner = model.forward(style)
    logits, att = captioner(features, caps, 0.0)
    loss = F.cross_entropy(logits.view(-1, V), caps.view(-1), ignore_index=0)
    model.zero_grad(set_to_none=True)
    loss.backward()
    out["tf/logits"], out["tf/attn"], out["tf/loss"] = logits, att, loss
    for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
        out["gen/" + k] = getattr(captioner.gru, k).detach().clone()
        out["tf/grad/captioner.gru." + k] = getattr(captioner.gru, k).grad.clone()
    for k, v in model.named_parameters():
        if k.startswith("image_encoder") or k.startswith("captioner.gru."):
            continue
        if v.grad is not None:
            out["tf/grad/" + k] = v.grad.clone()
        else:
            out.setdefault("tf/nograd", []).append(k)
    out["tf/nograd"] = np.array(out.get("tf/nograd", []))

    # greedy, test_hn.py path (cc_train_hypernet.py:228-230)
    np.random.seed(0)
    with torch.no_grad():
        captioner = model.forward(style)
        glogits, gatt = captioner(features, caps, 1.0)
    out["greedy/logits"], out["greedy/attn"] = glogits, gatt

    # B = 1 greedy_search with EOS stop (models/decoderlstm.py:138-175): takes features ALREADY through feature_fc
    with torch.no_grad():
        captioner = model.forward(style)
        for bi in range(B):
            fproj = captioner.feature_fc(features[bi:bi + 1])
            sent, wts = captioner.greedy_search(fproj, end_sentence=2, max_sentence=7)
            out[f"gs/{bi}/tokens"] = torch.tensor(sent)
            out[f"gs/{bi}/weights"] = torch.stack([w.reshape(-1) for w in wts], 0)

    # flow mode gradients (harness-side patch of set_all_parameters only)
    import hypernet_attention as ref_hna
    orig = ref_hna.set_all_parameters
    ref_hna.set_all_parameters = _flow_set_all_parameters
    try:
        torch.manual_seed(0)
        model2 = ref.HyperNetAttention(Fo, E, H, V, ref.vocab, cc=cc, hyper_emb=he)
        np.random.seed(0)
        captioner = model2.forward(style)
        logits2, _ = captioner(features, caps, 0.0)
        loss2 = F.cross_entropy(logits2.view(-1, V), caps.view(-1), ignore_index=0)
        loss2.backward()
        assert torch.equal(logits2, logits)
        for k, v in model2.named_parameters():
            if k.startswith("hn_") and v.grad is not None:
                out["flow/grad/" + k] = v.grad.clone()
    finally:
        ref_hna.set_all_parameters = orig
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(out))
    print(name, "loss", float(loss), "logits", tuple(logits.shape))


def case_pooled(ref, name, L, E=8, H=6, B=2, T=5, max_len=4):
    V = 9684  # later.py:449 hard-codes it
    torch.manual_seed(0)
    model = ref.HyperNetPooled(E, H, V, ref.vocab, num_layers=L)
    sd = _sd(model)
    g = torch.Generator().manual_seed(4321)
    pooled = torch.relu(torch.randn(B, 2048, generator=g))
    caps = torch.randint(7, 3000, (B, T), generator=g)
    caps[:, 0] = 1
    style = torch.randn(1, E, generator=g)
    out = {"pooled": pooled, "captions": caps, "style": style}
    out.update({"sd/" + k: v for k, v in sd.items()
                if not (k.startswith("captioner.lstm_cell.") or k.startswith("captioner.layers."))})

    captioner = model.forward(style)
    feats = model.image_encoder.fc(pooled)  # hypernet.py:46,134
    torch.manual_seed(1)
    h0 = torch.rand(B, H)
    torch.manual_seed(1)  # DecoderGRU.forward draws torch.rand(B,H) first (later.py:393)
    logits = captioner(feats, caps, True)
    loss = F.cross_entropy(logits.view(-1, V), caps.view(-1))  # hypernet.py:145, no ignore_index
    model.zero_grad(set_to_none=True)
    loss.backward()
    out["h0"], out["tf/logits"], out["tf/loss"] = h0, logits, loss
    cells = [captioner.lstm_cell] + (list(captioner.layers) if captioner.layers else [])
    for ci, cell in enumerate(cells):
        for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
            out[f"gen/{ci}/{k}"] = getattr(cell, k).detach().clone()
            out[f"tf/grad/gen/{ci}/{k}"] = getattr(cell, k).grad.clone()
    for k in ("image_encoder.fc.weight", "image_encoder.fc.bias", "captioner.embed.weight",
              "captioner.fc_out.weight", "captioner.fc_out.bias"):
        out["tf/grad/" + k] = dict(model.named_parameters())[k].grad.clone()

    with torch.no_grad():
        captioner = model.forward(style)
        torch.manual_seed(1)
        probs = captioner.infer(model.image_encoder.fc(pooled), max_len=max_len)
    out["infer/probs"] = probs

    import hypernet as ref_hn
    orig = ref_hn.set_all_parameters
    ref_hn.set_all_parameters = _flow_set_all_parameters
    try:
        torch.manual_seed(0)
        model2 = ref.HyperNetPooled(E, H, V, ref.vocab, num_layers=L)
        captioner = model2.forward(style)
        torch.manual_seed(1)
        logits2 = captioner(model2.image_encoder.fc(pooled), caps, True)
        loss2 = F.cross_entropy(logits2.view(-1, V), caps.view(-1))
        loss2.backward()
        assert torch.equal(logits2, logits)
        for k, v in model2.named_parameters():
            if k.startswith("hn_") and v.grad is not None:
                out["flow/grad/" + k] = v.grad.clone()
    finally:
        ref_hn.set_all_parameters = orig
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(out))
    print(name, "loss", float(loss), "logits", tuple(logits.shape))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_harness.load()
    g = torch.Generator().manual_seed(7)
    case_attention(ref, "attn_flickr", cc=False, style=torch.randn(1, 12, generator=g))
    onehot = torch.zeros(10)
    onehot[3] = 1.0
    case_attention(ref, "attn_cc", cc=True, style=onehot, he=10)
    case_pooled(ref, "pooled_l1", L=1)
    case_pooled(ref, "pooled_l2", L=2)


if __name__ == "__main__":
    main()
