"""Kernel-level checks of the many-style path: the grouped tcgen05 GEMM against fp64 per-group products, the many-G hypernet
(dense-GEMM layers) against the oracle hypernet, the fused grouped loss node against the unfused one."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from golden_util import rel_err, grad_close
from oracle import caption_hn_oracle as O


@pytest.mark.parametrize("G,B,T,E,Fd,H", [(3, 40, 5, 200, 200, 200), (17, 70, 3, 48, 40, 36), (100, 512, 4, 200, 200, 200)])
def test_grouped_gemms_match_per_group_products(G, B, T, E, Fd, H):
    from hypernet_image_captioning_b200 import ops
    from hypernet_image_captioning_b200.grouped import GroupPlan, _layout
    g = torch.Generator().manual_seed(G + B)
    groups = torch.randint(0, G, (B,), generator=g)
    plan = GroupPlan.get(groups, G, T, torch.device("cuda"))
    H3, o_hh, o_bi, o_bh, theta = _layout(E, Fd, H)
    gs = groups[plan.order.cpu()]
    Theta = (torch.randn(G, theta, generator=g) / 10).cuda()
    Xw = torch.randn(T * B, E, generator=g).cuda()
    b_ih = Theta[:, o_bi:o_bh].contiguous()
    Wsp = ops.split_bf16_batched(Theta, theta, E + Fd, G, H3, E + Fd)
    Xgm = ops.split_bf16_gather(Xw, plan.gm2tm, plan.R_gm)
    GIw = torch.full((T * B, H3), float("nan"), device="cuda")
    ops.gemm_tc_grouped(Xgm, False, Wsp, False, GIw, H3, plan.units_xproj(H3, E, 128), 128, bias=b_ih, rowmap=plan.gm2tm)
    W = Theta[:, :o_hh].reshape(G, H3, E + Fd).double().cpu()
    row_g = gs.repeat(T)                                            # group of time-major row t*B + b
    Xc, bc = Xw.double().cpu(), b_ih.double().cpu()
    ref = torch.zeros(T * B, H3, dtype=torch.float64)
    for gi in range(G):
        rows = (row_g == gi).nonzero().squeeze(1)
        ref[rows] = Xc[rows] @ W[gi][:, :E].t() + bc[gi]
    assert rel_err(GIw, ref) < 2e-5
    # dX
    dGI = torch.randn(T * B, H3, generator=g).cuda()
    dGI_gm = ops.split_bf16_gather(dGI, plan.gm2tm, plan.R_gm)
    dXw = torch.full((T * B, E), float("nan"), device="cuda")
    Wmn = ops.SplitOperand(Wsp.hi, Wsp.lo, E + Fd, G * H3, Wsp.ld, True)
    ops.gemm_tc_grouped(dGI_gm, False, Wmn, True, dXw, E, plan.units_dx(H3, E, 128), 128, rowmap=plan.gm2tm)
    ref = torch.zeros(T * B, E, dtype=torch.float64)
    for gi in range(G):
        rows = (row_g == gi).nonzero().squeeze(1)
        ref[rows] = dGI.double().cpu()[rows] @ W[gi][:, :E]
    assert rel_err(dXw, ref) < 2e-5
    # dW_ih, dW_hh, bias sums into dTheta
    XC = torch.randn(T * B, E + Fd, generator=g).cuda()
    Hp = torch.randn(T * B, H, generator=g).cuda()
    dGH = torch.randn(T * B, H3, generator=g).cuda()
    mn = lambda op: ops.SplitOperand(op.hi, op.lo, op.K, op.rows, op.ld, True)
    dTheta = torch.zeros(G, theta, device="cuda")
    ops.gemm_tc_grouped(mn(dGI_gm), True, mn(ops.split_bf16_gather(XC, plan.gm2tm, plan.R_gm)), True, dTheta, E + Fd,
                        plan.units_dw(H3, E + Fd, 128, theta, 0, E + Fd), 128)
    ops.gemm_tc_grouped(mn(ops.split_bf16_gather(dGH, plan.gm2tm, plan.R_gm)), True,
                        mn(ops.split_bf16_gather(Hp, plan.gm2tm, plan.R_gm)), True, dTheta, H,
                        plan.units_dw(H3, H, 128, theta, o_hh, H), 128)
    ops.group_colsum(dGI, plan.goff_dev, G, B, T, dTheta[:, o_bi:o_bh])
    ops.group_colsum(dGH, plan.goff_dev, G, B, T, dTheta[:, o_bh:])
    torch.cuda.synchronize()
    ref = torch.zeros(G, theta, dtype=torch.float64)
    dGIc, dGHc, XCc, Hpc = dGI.double().cpu(), dGH.double().cpu(), XC.double().cpu(), Hp.double().cpu()
    for gi in range(G):
        rows = (row_g == gi).nonzero().squeeze(1)
        ref[gi, :o_hh] = (dGIc[rows].t() @ XCc[rows]).reshape(-1)
        ref[gi, o_hh:o_bi] = (dGHc[rows].t() @ Hpc[rows]).reshape(-1)
        ref[gi, o_bi:o_bh] = dGIc[rows].sum(0)
        ref[gi, o_bh:] = dGHc[rows].sum(0)
    assert rel_err(dTheta, ref) < 2e-5


@pytest.mark.parametrize("G,he,E,Fo,H", [(12, 12, 12, 16, 20), (130, 130, 200, 200, 200)])
def test_many_group_hypernet_matches_oracle(G, he, E, Fo, H):
    """HyperNetThetaManyFn (dense tensor-core layers, G > 8) == oracle hypernet (hypernet_attention.py:111-118), values and
    every parameter gradient."""
    import hypernet_image_captioning_b200 as C
    p = O.init_params_attention(2048, Fo, E, H, 50, he, seed=3)
    g = torch.Generator().manual_seed(1)
    X = torch.randn(G, he, generator=g)
    wts = torch.randn(G, sum(int(np.prod(s)) for _, s in O.gru_param_shapes_attention(E, Fo, H)), generator=g)
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ref = torch.stack([O.hypernet_theta(pl, X[i:i + 1], 4).reshape(-1) for i in range(G)])
    (ref * wts).sum().backward()
    m = C.HyperNetAttention(Fo, E, H, 50, None, cc=True, hyper_emb=he)
    sd = m.state_dict(); sd.update(p); m.load_state_dict(sd)
    m = m.cuda()
    theta = m.generate_theta(X.cuda())
    assert theta.shape == ref.shape
    (theta * wts.cuda()).sum().backward()
    assert rel_err(theta, ref) < 1e-4
    for k, v in m.named_parameters():
        if k.startswith("hn_"):
            assert grad_close(v.grad, pl[k].grad, 1e-3), (k, rel_err(v.grad, pl[k].grad))


def test_grouped_fused_loss_matches_unfused():
    import hypernet_image_captioning_b200 as C
    B, T, Fo, E, H, V, G = 50, 6, 64, 48, 56, 500, 7
    p = O.init_params_attention(2048, Fo, E, H, V, E, seed=2)
    g = torch.Generator().manual_seed(8)
    feats = torch.randn(B, 49, 2048, generator=g).cuda()
    caps = O.synth_captions(B, T, V, g).cuda()
    styles = torch.randn(G, E, generator=g).cuda()
    groups = torch.randint(0, G, (B,), generator=g).cuda()
    res = []
    for fused in (False, True):
        m = C.HyperNetAttention(Fo, E, H, V, None)
        sd = m.state_dict(); sd.update(p); m.load_state_dict(sd)
        m = m.cuda()
        cap = m.forward_grouped(styles)
        np.random.seed(0)
        if fused:
            loss, logits, att = cap.forward_loss(feats, caps, 0.0, ignore_index=0, groups=groups)
        else:
            logits, att = cap(feats, caps, 0.0, groups=groups)
            loss = C.cross_entropy(logits, caps, 0)
        loss.backward()
        res.append((loss.item(), logits.detach(), {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}))
    assert abs(res[0][0] - res[1][0]) < 1e-6 * abs(res[0][0])
    assert torch.equal(res[0][1], res[1][1])
    for k, v in res[0][2].items():
        assert grad_close(res[1][2][k], v, 1e-3), k
