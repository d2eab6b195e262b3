"""bf16 mode: hypernet weights / gradients stored in bf16 (half the bytes of the dominant kernels), tensor-core products
with plain bf16 operands, fp32 accumulation and recurrent state.  Tolerances are stated here (measured: see asserts)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from golden_util import rel_err, grad_close
from oracle import caption_hn_oracle as O

BF16_TOL_LOGITS = 2e-2     # max|d| / max|ref| on logits vs the fp32 oracle; measured 6.2e-3 (pooled), 5.4e-3 (attention)
BF16_TOL_LOSS = 1e-2
BF16_TOL_GRAD = 6e-2


@pytest.fixture(autouse=True)
def _restore_precision():
    from hypernet_image_captioning_b200 import ops
    yield
    ops.set_precision("fp32")


@pytest.mark.parametrize("G,N,K", [(1, 97, 13), (1, 450, 450), (2, 1000, 1125), (3, 301, 843), (1, 700, 11250),
                                   (4, 64, 8437), (1, 5000, 480), (5, 77, 6), (1, 5, 1)])
@pytest.mark.parametrize("act", [0, 1])
def test_rows_linear_bf16_weights(G, N, K, act):
    """The bf16 streaming kernels compute exactly act(A @ float(W_bf16)^T + b) in fp32 and round dW to bf16."""
    from hypernet_image_captioning_b200 import ops
    g = torch.Generator().manual_seed(G * 1000 + N + K)
    Wb = torch.randn(N, K, generator=g).bfloat16()
    W = Wb.double().requires_grad_(True)
    b = torch.randn(N, generator=g).double().requires_grad_(True)
    A = torch.randn(G, K, generator=g).double().requires_grad_(True)
    Yr = F.linear(A, W, b)
    if act:
        Yr = F.leaky_relu(Yr, 0.01)
    dY = torch.randn(G, N, generator=g).double()
    Yr.backward(dY)
    Wc, Ac = Wb.cuda(), A.detach().float().cuda()
    Y = ops.rows_linear_fwd(Wc, b.detach().float().cuda(), Ac, act)
    assert rel_err(Y, Yr) < 2e-6
    dW, db, dA = ops.rows_linear_bwd(Wc, Ac, Y, dY.float().cuda(), act)
    assert dW.dtype == torch.bfloat16
    assert rel_err(dW.float(), W.grad) < 5e-3          # one bf16 rounding of the fp32 result
    assert rel_err(db, b.grad) < 2e-6 and rel_err(dA, A.grad) < 1e-5


def test_pooled_bf16_mode_within_stated_tolerance():
    import hypernet_image_captioning_b200 as C
    B, T, E, H, V = 64, 20, 200, 150, 2000
    p = O.init_params_pooled(2048, E, H, V, seed=5)
    g = torch.Generator().manual_seed(99)
    pooled = torch.relu(torch.randn(B, 2048, generator=g))
    caps = O.synth_captions(B, T, V, g)
    style, h0 = torch.randn(1, E, generator=g), torch.rand(B, H, generator=g)
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lr, _, _ = O.path_pooled(pl, style, pooled, caps, h0)
    loss_ref = O.caption_loss(lr, caps, None)
    loss_ref.backward()
    m = C.HyperNetPooled(E, H, V, None)
    sd = m.state_dict(); sd.update(p); m.load_state_dict(sd)
    m = m.cuda().set_precision("bf16")
    assert m.hn_heads[0][2].weight.dtype == torch.bfloat16
    cap = m.forward(style.cuda())
    loss, logits = cap.forward_loss(m.image_encoder(pooled.cuda()), caps.cuda(), h0=h0.cuda())
    loss.backward()
    e_logits, e_loss = rel_err(logits, lr), abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())
    print(f"bf16 mode: logits rel err {e_logits:.2e}, loss rel err {e_loss:.2e}")
    assert e_logits < BF16_TOL_LOGITS and e_loss < BF16_TOL_LOSS
    named = dict(m.named_parameters())
    assert named["hn_heads.0.2.weight"].grad.dtype == torch.bfloat16
    for k in ("hn_heads.0.2.weight", "hn_heads.1.2.weight", "captioner.fc_out.weight", "captioner.embed.weight"):
        assert grad_close(named[k].grad.float(), pl[k].grad, BF16_TOL_GRAD), k


def test_attention_bf16_mode_within_stated_tolerance():
    import numpy as np
    import hypernet_image_captioning_b200 as C
    B, T, Fo, E, H, V = 32, 12, 200, 200, 200, 1500
    p = O.init_params_attention(2048, Fo, E, H, V, E, seed=7)
    g = torch.Generator().manual_seed(17)
    feats = torch.randn(B, 49, 2048, generator=g)
    caps = O.synth_captions(B, T, V, g)
    style = torch.randn(1, E, generator=g)
    with torch.no_grad():
        lr, ar, _, _ = O.path_attention(p, style, feats, caps, 0.0, np.random.RandomState(0))
    m = C.HyperNetAttention(Fo, E, H, V, None)
    sd = m.state_dict(); sd.update(p); m.load_state_dict(sd)
    m = m.cuda().set_precision("bf16")
    loss, logits, att = m.forward(style.cuda()).forward_loss(feats.cuda(), caps.cuda(), 0.0, ignore_index=0)
    loss.backward()
    e = rel_err(logits, lr)
    print(f"attention bf16 mode: logits rel err {e:.2e}, attn rel err {rel_err(att, ar):.2e}")
    assert e < BF16_TOL_LOGITS and rel_err(att, ar) < BF16_TOL_LOGITS
    assert torch.isfinite(dict(m.named_parameters())["hn_heads.0.2.weight"].grad.float()).all()
