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
