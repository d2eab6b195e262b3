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


def test_fused_adam_bf16_parameters_with_fp32_master():
    """bf16 mode is trainable: FusedAdam keeps an fp32 master copy + fp32 moments for bf16 parameters, reads the bf16
    gradient, writes the master weight and its bf16 rounding.  Reference: torch.optim.Adam + clip_grad_norm_ on fp32 copies
    fed with the same (bf16-valued) gradients; the bf16 parameter must equal bf16(master) after every step."""
    import hypernet_image_captioning_b200 as C
    g = torch.Generator().manual_seed(0)
    shapes = [(257, 33), (1000,), (64, 128)]
    ps = [torch.nn.Parameter((torch.randn(*s, generator=g) * 0.1).cuda().to(torch.bfloat16)) for s in shapes]
    refs = [torch.nn.Parameter(p.detach().float().clone()) for p in ps]
    opt = C.FusedAdam(ps, lr=1e-2, max_grad_norm=1.0)
    ref_opt = torch.optim.Adam(refs, lr=1e-2)
    for step in range(4):
        for p, r in zip(ps, refs):
            gr = (torch.randn(*p.shape, generator=g) * (3.0 if step == 1 else 0.3)).cuda().to(torch.bfloat16)
            p.grad = gr
            r.grad = gr.float().clone()
        torch.nn.utils.clip_grad_norm_(refs, 1.0)
        ref_opt.step()
        opt.step()
        for p, r in zip(ps, refs):
            master = opt.state[p]["master"]
            assert (master - r.detach()).abs().max().item() <= 2e-6 * max(1.0, r.detach().abs().max().item()), step
            assert torch.equal(p.detach(), master.to(torch.bfloat16))


def test_bf16_mode_training_step_reduces_loss():
    """End to end in bf16 mode: forward + backward + FusedAdam on the bf16 hypernet parameters lowers the caption loss."""
    import hypernet_image_captioning_b200 as C
    from hypernet_image_captioning_b200 import ops
    from hypernet_image_captioning_b200.synth import synth_captions
    torch.manual_seed(0)
    with torch.device("cuda"):
        m = C.HyperNetPooled(48, 40, 500, None)
    m.set_precision("bf16")
    try:
        g = torch.Generator().manual_seed(1)
        pooled = torch.relu(torch.randn(32, 2048, generator=g)).cuda()
        caps = synth_captions(32, 8, 500, g).cuda()
        h0 = torch.rand(32, 40, generator=g).cuda()
        opt = C.FusedAdam([p for n, p in m.named_parameters() if not n.startswith("captioner.lstm_cell")], lr=2e-3,
                          max_grad_norm=5.0)
        losses = []
        for _ in range(8):
            m.zero_grad(set_to_none=True)
            cap = m.forward(m.captioner.embed.weight[4:5])
            loss, _ = cap.forward_loss(m.image_encoder(pooled), caps, h0=h0)
            loss.backward()
            opt.step()
            losses.append(loss.item())
        assert all(p.dtype == torch.bfloat16 for p in m.hn_heads.parameters())
        assert losses[-1] < losses[0] - 0.05, losses
    finally:
        ops.set_precision("fp32")
