"""Edge cases of the hot path through the drop-in API: ragged / minimal shapes, all-padding batches, odd sizes."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from golden_util import rel_err, grad_close
from oracle import caption_hn_oracle as O


def _load(m, p):
    sd = m.state_dict(); sd.update(p); m.load_state_dict(sd)
    return m.cuda()


@pytest.mark.parametrize("B,T", [(1, 1), (1, 2), (3, 3), (5, 20), (33, 2), (130, 5)])
def test_attention_small_and_ragged_shapes(B, T):
    """T = 1 and T = 2 exercise the zero-input steps only (models/decoderlstm.py:82-88); B not a multiple of the row tile."""
    import hypernet_image_captioning_b200 as C
    Fo, E, H, V = 16, 12, 20, 60
    p = O.init_params_attention(2048, Fo, E, H, V, E, seed=11)
    g = torch.Generator().manual_seed(B * 7 + T)
    feats = torch.randn(B, 49, 2048, generator=g)
    caps = torch.randint(4, V, (B, T), generator=g)
    caps[:, 0] = 1
    style = torch.randn(1, E, generator=g)
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lr, ar, _, _ = O.path_attention(pl, style, feats, caps, 0.0, np.random.RandomState(0))
    O.caption_loss(lr, caps, 0).backward()
    m = _load(C.HyperNetAttention(Fo, E, H, V, None), p)
    logits, att = m.forward(style.cuda())(feats.cuda(), caps.cuda(), 0.0)
    C.cross_entropy(logits, caps.cuda(), 0).backward()
    assert rel_err(logits, lr) < 1e-4 and rel_err(att, ar) < 1e-4
    for k in ("hn_heads.1.2.weight", "captioner.embed.weight", "captioner.feature_fc.0.weight", "captioner.init_h.bias"):
        assert grad_close(dict(m.named_parameters())[k].grad, pl[k].grad, 1e-3), k
    with torch.no_grad():
        gl_ref, _, _, _ = O.path_attention(p, style, feats, caps, 1.0, np.random.RandomState(0))
        gl, _ = m.forward(style.cuda())(feats.cuda(), caps.cuda(), 1.0)
    assert torch.equal(gl.argmax(-1).cpu(), gl_ref.argmax(-1))


@pytest.mark.parametrize("B,T,L", [(1, 1, 1), (2, 2, 2), (7, 3, 1), (19, 20, 1), (9, 4, 3)])
def test_pooled_small_and_ragged_shapes(B, T, L):
    import hypernet_image_captioning_b200 as C
    E, H, V = 16, 12, 97
    p = O.init_params_pooled(2048, E, H, V, L=L, seed=13)
    g = torch.Generator().manual_seed(B * 5 + T)
    pooled = torch.relu(torch.randn(B, 2048, generator=g))
    caps = torch.randint(4, V, (B, T), generator=g)
    style, h0 = torch.randn(1, E, generator=g), torch.rand(B, H, generator=g)
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lr, _, _ = O.path_pooled(pl, style, pooled, caps, h0, L=L)
    O.caption_loss(lr, caps, None).backward()
    m = _load(C.HyperNetPooled(E, H, V, None, num_layers=L), p)
    cap = m.forward(style.cuda())
    loss, logits = cap.forward_loss(m.image_encoder(pooled.cuda()), caps.cuda(), h0=h0.cuda())
    loss.backward()
    assert rel_err(logits, lr) < 1e-4
    for k, v in m.named_parameters():
        if k.startswith("captioner.lstm_cell.") or k.startswith("captioner.layers."):
            continue
        if pl[k].grad is None:   # T = 1: the embedding table is never read by the reference -> no gradient at all
            assert v.grad is None or float(v.grad.abs().max()) == 0.0, k
            continue
        assert grad_close(v.grad, pl[k].grad, 1e-3), k


def test_all_padding_batch_gives_nan_loss_like_reference():
    """F.cross_entropy(ignore_index=0) over targets that are all <pad> is 0/0 = NaN (cc_train_hypernet.py:153)."""
    import hypernet_image_captioning_b200 as C
    x = torch.randn(6, 11).cuda()
    t = torch.zeros(6, dtype=torch.long).cuda()
    assert torch.isnan(C.cross_entropy(x, t, 0)).item()
    assert torch.isnan(F.cross_entropy(x.cpu(), t.cpu(), ignore_index=0)).item()


def test_padded_rows_do_not_contribute():
    """Rows that are entirely <pad> after position 0 change neither the loss nor the gradients of the other rows."""
    import hypernet_image_captioning_b200 as C
    Fo, E, H, V, B, T = 16, 12, 20, 60, 4, 6
    p = O.init_params_attention(2048, Fo, E, H, V, E, seed=21)
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(B, 49, 2048, generator=g)
    caps = O.synth_captions(B, T, V, g)
    caps[3, :] = 0
    style = torch.randn(1, E, generator=g)
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lr, _, _, _ = O.path_attention(pl, style, feats, caps, 0.0, np.random.RandomState(0))
    loss_ref = O.caption_loss(lr, caps, 0)
    loss_ref.backward()
    m = _load(C.HyperNetAttention(Fo, E, H, V, None), p)
    loss, logits, _ = m.forward(style.cuda()).forward_loss(feats.cuda(), caps.cuda(), 0.0, ignore_index=0)
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) < 1e-4 * abs(loss_ref.item())
    assert grad_close(dict(m.named_parameters())["captioner.fc.weight"].grad, pl["captioner.fc.weight"].grad, 1e-3)


def test_hypernet_cc_embedding_front_ends():
    """Style vectors of the CC front-ends reach HyperNet.forward as 1-D [he] tensors (cc_train_hypernet.py:137-149):
    one-hot row (he = #domains) and nn.Embedding row (he = 10)."""
    import hypernet_image_captioning_b200 as C
    for he, style in ((37, torch.eye(37)[5]), (10, torch.randn(10, generator=torch.Generator().manual_seed(1)))):
        Fo, E, H, V = 16, 12, 20, 60
        p = O.init_params_attention(2048, Fo, E, H, V, he, seed=3)
        m = _load(C.HyperNetAttention(Fo, E, H, V, None, cc=True, hyper_emb=he), p)
        theta_ref = O.hypernet_theta(p, style, 4)
        cap = m.forward(style.cuda())
        got = torch.cat([getattr(cap.gru, k).detach().flatten() for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")])
        assert rel_err(got, theta_ref) < 1e-5


def test_sweep_dims_attention_hidden_256():
    """BASELINE config 5 direction: larger hidden size (H = 256, E = F = 200): exercises wider recurrence tiles."""
    import hypernet_image_captioning_b200 as C
    B, T, Fo, E, H, V = 24, 16, 200, 200, 256, 500
    p = O.init_params_attention(2048, Fo, E, H, V, E, seed=31)
    g = torch.Generator().manual_seed(2)
    feats = torch.randn(B, 49, 2048, generator=g)
    caps = O.synth_captions(B, T, V, g)
    style = torch.randn(1, E, generator=g)
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lr, ar, _, _ = O.path_attention(pl, style, feats, caps, 0.0, np.random.RandomState(0))
    O.caption_loss(lr, caps, 0).backward()
    m = _load(C.HyperNetAttention(Fo, E, H, V, None), p)
    loss, logits, att = m.forward(style.cuda()).forward_loss(feats.cuda(), caps.cuda(), 0.0, ignore_index=0)
    loss.backward()
    assert rel_err(logits, lr) < 1e-4 and rel_err(att, ar) < 1e-4
    for k in ("hn_heads.0.2.weight", "hn_heads.1.2.weight", "captioner.attention.U_a.weight", "captioner.fc.weight"):
        assert grad_close(dict(m.named_parameters())[k].grad, pl[k].grad, 1e-3), k


@pytest.mark.parametrize("H,E", [(96, 32), (200, 16), (33, 24)])
def test_sweep_dims_pooled_cluster_sizes(H, E):
    """Hidden sizes that make the weights-resident GRU kernel pick cluster sizes 2 / 4 (and odd strides)."""
    import hypernet_image_captioning_b200 as C
    from hypernet_image_captioning_b200 import ops
    B, T, V = 40, 24, 300
    assert ops.gru_cluster_size(H) in (2, 4, 8)
    p = O.init_params_pooled(2048, E, H, V, seed=41)
    g = torch.Generator().manual_seed(3)
    pooled = torch.relu(torch.randn(B, 2048, generator=g))
    caps = O.synth_captions(B, T, V, g)
    style, h0 = torch.randn(1, E, generator=g), torch.rand(B, H, generator=g)
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lr, _, _ = O.path_pooled(pl, style, pooled, caps, h0)
    O.caption_loss(lr, caps, None).backward()
    m = _load(C.HyperNetPooled(E, H, V, None), p)
    cap = m.forward(style.cuda())
    loss, logits = cap.forward_loss(m.image_encoder(pooled.cuda()), caps.cuda(), h0=h0.cuda())
    loss.backward()
    assert rel_err(logits, lr) < 1e-4
    for k in ("hn_heads.1.2.weight", "captioner.embed.weight", "captioner.fc_out.weight", "image_encoder.fc.weight"):
        assert grad_close(dict(m.named_parameters())[k].grad, pl[k].grad, 1e-3), k


# ---- BASELINE configs[4] sweep corners: hidden 512-1024, caption length 64, batch 4096 -----------------------------
@pytest.mark.parametrize("H,Fo,E,B,T", [(512, 40, 24, 6, 4), (208, 208, 16, 5, 3), (64, 32, 16, 4, 64)])
def test_sweep_corners_attention(H, Fo, E, B, T):
    """Hidden sizes beyond the step-split kernels (they fall back to the persistent streaming kernels), the largest
    covered size (H = F = 208) and the longest caption of the sweep, against the oracle: logits, loss, gradients, greedy."""
    import hypernet_image_captioning_b200 as C
    V = 70
    p = O.init_params_attention(2048, Fo, E, H, V, E, seed=5)
    g = torch.Generator().manual_seed(H + T)
    feats = torch.randn(B, 49, 2048, generator=g)
    caps = O.synth_captions(B, T, V, g)
    style = torch.randn(1, E, generator=g)
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lr, ar, _, _ = O.path_attention(pl, style, feats, caps, 0.0, np.random.RandomState(0))
    loss_ref = O.caption_loss(lr, caps, 0)
    loss_ref.backward()
    m = _load(C.HyperNetAttention(Fo, E, H, V, None), p)
    loss, logits, att = m.forward(style.cuda()).forward_loss(feats.cuda(), caps.cuda(), 0.0, ignore_index=0)
    loss.backward()
    assert rel_err(logits, lr) < 1e-4 and rel_err(att, ar) < 1e-4
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 1e-4 * abs(float(loss_ref.detach()))
    for k in ("hn_heads.0.2.weight", "hn_heads.1.2.weight", "captioner.attention.U_a.weight", "captioner.attention.v_a.weight",
              "captioner.embed.weight", "captioner.feature_fc.2.weight", "captioner.init_h.weight"):
        assert grad_close(dict(m.named_parameters())[k].grad, pl[k].grad, 2e-3), k
    with torch.no_grad():
        gl_ref, _, _, _ = O.path_attention(p, style, feats, caps, 1.0, np.random.RandomState(0))
        gl, _ = m.forward(style.cuda())(feats.cuda(), caps.cuda(), 1.0)
    assert torch.equal(gl.argmax(-1).cpu(), gl_ref.argmax(-1))


def test_sweep_corner_hidden_1024_recurrence():
    """H = 1024 (the sweep's maximum; its hypernet heads alone are 80 GB, too large for the CPU oracle): the recurrence
    kernels against the oracle's per-step functions (models/attention.py:33-45 + GRUCell) on generated-weight-sized
    random matrices, forward and BPTT."""
    from hypernet_image_captioning_b200 import ops
    B, T, P, H, Fo, E = 3, 3, 49, 1024, 24, 16
    g = torch.Generator().manual_seed(1024)
    r = lambda *s: torch.randn(*s, generator=g)
    f, K = r(B, P, Fo) * 0.5, r(B, P, H) * 0.5
    Ua, bu, va, bv = r(H, H) / H ** 0.5, r(H) * 0.1, r(H) * 0.05, r(1)
    W_ih, W_hh = r(3 * H, E + Fo) / (E + Fo) ** 0.5, r(3 * H, H) / H ** 0.5
    b_ih, b_hh, h0 = r(3 * H) * 0.1, r(3 * H) * 0.1, r(B, H) * 0.5
    xw = r(T, B, E) * 0.5
    leaves = [t_.clone().requires_grad_(True) for t_ in (K, Ua, va, W_ih, W_hh, h0)]
    Kl, Ual, val, Wil, Whl, h = leaves
    hs, als = [], []
    for t in range(T):                                     # oracle step: bahdanau + gru_cell with explicit tensors
        u = F.linear(h, Ual, bu)
        s = torch.tanh(Kl + u[:, None, :]) @ val + bv
        al = torch.softmax(s, dim=1)
        ctx = (al[:, :, None] * f).sum(1)
        h = O.gru_cell(torch.cat([xw[t], ctx], 1), h, Wil, Whl, b_ih, b_hh)
        hs.append(h); als.append(al)
    Href, Aref = torch.stack(hs, 1), torch.stack(als, 1)
    dH = r(B, T, H) * 0.3
    (Href * dH).sum().backward()
    c = lambda t_: t_.detach().cuda().contiguous()
    GIw = F.linear(xw.reshape(T * B, E), W_ih[:, :E], b_ih).cuda()
    lw = ops.AttGruWeights(c(W_ih), c(W_hh), c(Ua), E, P)
    assert lw.pack is None                                  # H = 1024 is beyond the step-split kernels: persistent path
    Hall = torch.empty(T + 1, B, H).cuda(); Hall[0] = c(h0)
    Hbm, attn = torch.empty(B, T, H).cuda(), torch.empty(B, T, P).cuda()
    XC, saved = torch.zeros(T * B, E + Fo).cuda(), torch.empty(5, T, B, H).cuda()
    ops.attgru_fwd(c(K), c(f), GIw, lw, c(bu), c(va), c(bv), c(b_hh), Hall, Hbm, attn, XC, E, saved, 0, T)
    assert rel_err(Hbm, Href) < 1e-4 and rel_err(attn, Aref) < 1e-4
    dGI, dGH, dU, dCTX, dK, dva, dbv, dh0 = ops.attgru_bwd(c(dH), None, c(K), c(f), attn, saved, Hall, c(Ua), c(va),
                                                          c(W_ih), c(W_hh), E)
    assert grad_close(dK, Kl.grad, 1e-3) and grad_close(dh0, leaves[5].grad, 1e-3)
    assert grad_close(dva, val.grad, 1e-3)
    Hprev = Hall[:-1].reshape(T * B, H)
    assert grad_close(dGH.t() @ Hprev, Whl.grad, 1e-3)     # dW_hh = sum_t dgh_t^T h_{t-1}
    assert grad_close(dU.t() @ Hprev, Ual.grad, 1e-3)


def test_sweep_corner_pooled_caption_length_64():
    import hypernet_image_captioning_b200 as C
    E, H, V, B, T = 16, 48, 200, 5, 64
    p = O.init_params_pooled(2048, E, H, V, seed=43)
    g = torch.Generator().manual_seed(6)
    pooled = torch.relu(torch.randn(B, 2048, generator=g))
    caps = O.synth_captions(B, T, V, g)
    style, h0 = torch.randn(1, E, generator=g), torch.rand(B, H, generator=g)
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lr, _, _ = O.path_pooled(pl, style, pooled, caps, h0)
    O.caption_loss(lr, caps, None).backward()
    m = _load(C.HyperNetPooled(E, H, V, None), p)
    loss, logits = m.forward(style.cuda()).forward_loss(m.image_encoder(pooled.cuda()), caps.cuda(), h0=h0.cuda())
    loss.backward()
    assert rel_err(logits, lr) < 1e-4
    for k in ("hn_heads.1.2.weight", "captioner.embed.weight", "captioner.fc_out.weight"):
        assert grad_close(dict(m.named_parameters())[k].grad, pl[k].grad, 1e-3), k


@pytest.mark.parametrize("variant", ["attention", "pooled"])
def test_sweep_corner_batch_4096_is_batch_invariant(variant):
    """Batch 4096 (the sweep's maximum; too large for the CPU oracle): every row's logits must equal the same row
    decoded in a batch of 64 -- rows are independent given (theta, shared parameters) -- and the loss must be finite."""
    import hypernet_image_captioning_b200 as C
    B, Bs, T, V = 4096, 64, 6, 80
    g = torch.Generator().manual_seed(9)
    caps = O.synth_captions(B, T, V, g).cuda()
    with torch.no_grad():
        if variant == "attention":
            Fo, E, H = 16, 12, 20
            m = _load(C.HyperNetAttention(Fo, E, H, V, None), O.init_params_attention(2048, Fo, E, H, V, E, seed=3))
            feats = torch.randn(B, 49, 2048, generator=g).cuda()
            cap = m.forward(torch.randn(1, E, generator=g).cuda())
            big, _ = cap(feats, caps, 0.0)
            small, _ = cap(feats[1000:1000 + Bs].contiguous(), caps[1000:1000 + Bs].contiguous(), 0.0)
            loss = C.cross_entropy(big, caps, 0)
        else:
            E, H = 16, 24
            m = _load(C.HyperNetPooled(E, H, V, None), O.init_params_pooled(2048, E, H, V, seed=3))
            pooled = torch.relu(torch.randn(B, 2048, generator=g)).cuda()
            h0 = torch.rand(B, H, generator=g).cuda()
            cap = m.forward(torch.randn(1, E, generator=g).cuda())
            big = cap.forward_loss(m.image_encoder(pooled), caps, h0=h0)[1]
            small = cap.forward_loss(m.image_encoder(pooled[1000:1000 + Bs].contiguous()), caps[1000:1000 + Bs].contiguous(),
                                     h0=h0[1000:1000 + Bs].contiguous())[1]
            loss = C.cross_entropy(big, caps, None)
    assert rel_err(big[1000:1000 + Bs], small) < 2e-5
    assert torch.isfinite(loss)


@pytest.mark.parametrize("he", [100, 150])
def test_conceptual_captions_one_hot_domains(he):
    """BASELINE configs[3] / SURVEY C4: cc=True with a one-hot domain vector of 100 / 150 domains as the hypernet input
    (cc_train_hypernet.py:86-89,141-144: 1-D [he] float), against the oracle: logits, loss, gradients, greedy tokens."""
    import hypernet_image_captioning_b200 as C
    Fo, E, H, V, B, T = 16, 12, 56, 60, 7, 5          # 3H = 168 >= he: the reference's sane head branch (hypernet_attention.py:85-99)
    p = O.init_params_attention(2048, Fo, E, H, V, he, seed=he)
    g = torch.Generator().manual_seed(he)
    feats = torch.randn(B, 49, 2048, generator=g)
    caps = O.synth_captions(B, T, V, g)
    style = torch.zeros(he)
    style[he // 3] = 1.0
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lr, ar, _, _ = O.path_attention(pl, style, feats, caps, 0.0, np.random.RandomState(0))
    loss_ref = O.caption_loss(lr, caps, 0)
    loss_ref.backward()
    m = _load(C.HyperNetAttention(Fo, E, H, V, None, cc=True, hyper_emb=he), p)
    loss, logits, att = m.forward(style.cuda()).forward_loss(feats.cuda(), caps.cuda(), 0.0, ignore_index=0)
    loss.backward()
    assert rel_err(logits, lr) < 1e-4 and rel_err(att, ar) < 1e-4
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 1e-4 * abs(float(loss_ref.detach()))
    for k in ("hn_base.0.weight", "hn_heads.0.0.weight", "hn_heads.0.2.weight", "hn_heads.3.2.bias", "captioner.embed.weight"):
        assert grad_close(dict(m.named_parameters())[k].grad, pl[k].grad, 1e-3), k
    with torch.no_grad():
        gl_ref, _, _, _ = O.path_attention(p, style, feats, caps, 1.0, np.random.RandomState(0))
        gl, _ = m.forward(style.cuda())(feats.cuda(), caps.cuda(), 1.0)
    assert torch.equal(gl.argmax(-1).cpu(), gl_ref.argmax(-1))
