"""Subprocess body of tests/test_dropin_reference_scripts.py: imports the UNEDITED reference scripts on top of dropin/.

Runs in its own interpreter because it rewires ``sys.path`` / ``sys.modules`` (stub third-party packages from
oracle/ref_harness.py, dropin/ ahead of the reference checkout).  Prints one JSON line.
"""
import json
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(mode):
    import numpy as np
    import torch
    from oracle import ref_harness as RH
    RH.install_stubs()
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    import run_reference
    run_reference.install(RH.REFERENCE_DIR)
    os.chdir(RH.REFERENCE_DIR)
    import build_vocab
    sys.modules["__main__"].Vocab = build_vocab.Vocab            # data/vocab.pkl was pickled from __main__

    import cc_train_hypernet, train_hyper_combine, test_hn      # noqa: E401  (cc_train_hypernet.py:10-28 etc.)
    import hypernet_image_captioning_b200 as C
    out = {
        "HyperNet_cc": cc_train_hypernet.HyperNet is C.HyperNetAttention,
        "HyperNet_combine": train_hyper_combine.HyperNet is C.HyperNetAttention,
        "HyperNet_test_hn": test_hn.HyperNet is C.HyperNetAttention,
        "AttentionGru": cc_train_hypernet.AttentionGru is C.AttentionGru,
        "EncoderCNN_module": cc_train_hypernet.EncoderCNN.__module__,
        "EncoderCNN_file": os.path.realpath(sys.modules["models.encoder"].__file__),
    }
    with open("data/vocab.pkl", "rb") as fh:
        vocab = pickle.load(fh)
    V = len(vocab)
    domains = [f"domain{i}.com" for i in range(5)]
    torch.manual_seed(0)
    model = cc_train_hypernet.HyperNetCC(200, 200, 200, V, vocab, domains)      # one-hot: he = #domains
    out["hypernet_type"] = type(model.hypernet).__name__
    out["hypernet_is_ours"] = type(model.hypernet) is C.HyperNetAttention
    out["captioner_is_ours"] = type(model.hypernet.captioner) is C.AttentionGru
    out["he"] = model.hypernet.hn_base[0].in_features

    # the reference's metric step needs nltk / datasets metrics that the image lacks: stub it (ref_harness-style)
    cc_train_hypernet.metric_score = lambda *a, **k: (0.0,) * 7
    g = torch.Generator().manual_seed(7)
    B, T = 6, 9
    feats = torch.randn(B, 49, 2048, generator=g)
    caps = torch.randint(7, V, (B, T), generator=g)
    caps[:, 0], caps[:, -1] = 1, 2
    caps[0, 5:] = 0
    caps[0, 4] = 2
    batch_domains = [domains[3]] * B
    if mode == "cpu":
        # no GPU here: the unedited training_step must reach OUR kernels and fail loudly (no CPU fallback)
        try:
            model.training_step((feats, caps, None, batch_domains), 0)
            out["training_step"] = "ran on CPU (a CPU fallback exists: wrong)"
        except C._cabi.CaphnError as e:
            out["training_step"] = "CaphnError"
            out["error"] = str(e)
    else:
        from oracle import caption_hn_oracle as O
        dev = torch.device("cuda", 0)
        model = model.to(dev)
        type(model).device = property(lambda self: dev)
        np.random.seed(0)
        loss = model.training_step((feats.to(dev), caps.to(dev), None, batch_domains), 0)
        loss.backward()
        p = {k[len("hypernet."):]: v.detach().cpu() for k, v in model.state_dict().items() if k.startswith("hypernet.")}
        style = torch.nn.functional.one_hot(torch.tensor(3), len(domains)).float()
        np.random.seed(0)
        logits = O.path_attention(p, style, feats, caps)[0]
        ref = O.caption_loss(logits, caps, 0)
        out["loss"], out["oracle_loss"] = float(loss), float(ref)
        out["head_grad_set"] = model.hypernet.hn_heads[0][2].weight.grad is not None
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "cpu")
