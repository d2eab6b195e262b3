"""Host logic of the many-style path (hypernet_image_captioning_b200/grouped.py GroupPlan) on the CPU: the permutation,
the group-major row map, the recurrence tile tables and the grouped-GEMM unit tables are checked by EMULATING the grouped
tensor-core launch (include/caphn_b200.h caphn_gemm_tc_grouped: unit records, K-major / MN-major operand addressing,
zero fill outside the arrays, row map, C offsets) in numpy and comparing with the per-group products computed directly."""
import numpy as np
import pytest
import torch

from hypernet_image_captioning_b200.grouped import GroupPlan, _layout


def _emulate(A, a_mn, Bm, b_mn, C, ldc, units, BN, bias=None, rowmap=None):
    """numpy model of caphn_gemm_tc_grouped on fp64 'operand arrays' A, Bm (2-D, already padded like the bf16 buffers)."""
    Cf = C.reshape(-1)

    def tile(X, mn, row0, nrows, k0, nk):
        out = np.zeros((nrows, nk))
        for r in range(nrows):
            for k in range(nk):
                i, j = (k0 + k, row0 + r) if mn else (row0 + r, k0 + k)
                if 0 <= i < X.shape[0] and 0 <= j < X.shape[1]:
                    out[r, k] = X[i, j]
        return out

    for u in units:
        a_row, b_row, ka0, kb0, nkb, mv, nv, boff, map0, clo, chi, _ = [int(v) for v in u]
        coff = (chi << 32) | (clo & 0xFFFFFFFF)
        At = tile(A, a_mn, a_row, 128, ka0, nkb * 64)
        Bt = tile(Bm, b_mn, b_row, BN, kb0, nkb * 64)
        acc = At @ Bt.T
        for r in range(mv):
            orow = r if rowmap is None else int(rowmap[map0 + r])
            if orow < 0:
                continue
            for c in range(nv):
                Cf[coff + orow * ldc + c] = acc[r, c] + (bias[boff + c] if bias is not None else 0.0)


def _gather(src, rowmap, Kp):
    out = np.zeros((len(rowmap), Kp))
    for i, r in enumerate(rowmap):
        if r >= 0:
            out[i, :src.shape[1]] = src[r]
    return out


@pytest.mark.parametrize("G,B,T,seed", [(3, 7, 3, 0), (5, 9, 2, 1), (4, 6, 4, 2)])
def test_group_plan_tables_reproduce_per_group_products(G, B, T, seed):
    rng = np.random.RandomState(seed)
    E, Fd, H = 5, 3, 4
    H3, o_hh, o_bi, o_bh, theta = _layout(E, Fd, H)
    groups = rng.randint(0, G, size=B)
    if G == 4:
        groups[groups == 2] = 1                                  # an absent group
    plan = GroupPlan(groups.astype(np.int64), G, T, "cpu")
    order = plan.order.numpy()
    assert (np.sort(groups) == groups[order]).all() and (order[plan.inv.numpy()] == np.arange(B)).all()
    gs = groups[order]                                           # group of sorted row b
    # tile tables: every sorted row exactly once, no tile straddles a group
    for tiles, nb in ((plan.tiles_fwd.numpy(), 64), (plan.tiles_bwd.numpy(), 32)):
        seen = np.zeros(B, int)
        for r0, n, g, _ in tiles:
            assert 0 < n <= nb and (gs[r0:r0 + n] == g).all()
            seen[r0:r0 + n] += 1
        assert (seen == 1).all()
    # group-major map: each time-major row exactly once, inside its group's block, blocks padded to 64
    gm2tm = plan.gm2tm.numpy()
    assert plan.R_gm % 64 == 0 and sorted(gm2tm[gm2tm >= 0].tolist()) == list(range(T * B))
    for i, r in enumerate(gm2tm):
        if r >= 0:
            g = gs[r % B]
            assert plan.gm_off[g] <= i < plan.gm_off[g] + plan.gm_cnt[g]
    Theta = rng.randn(G, theta)
    W_ih = [Theta[g, :o_hh].reshape(H3, E + Fd) for g in range(G)]
    b_ih = Theta[:, o_bi:o_bh].copy()
    Xw = rng.randn(T * B, E)                                      # time-major, sorted batch
    # ---- x-projection: A = X group-major (K-major), B = W_ih of all groups [G*H3, Kp], scatter through gm2tm ----
    Kp_e, Kp_ef = 64, 64
    Xgm = _gather(Xw, gm2tm, Kp_e)
    Wsp = np.zeros((G * H3, Kp_ef))
    for g in range(G):
        Wsp[g * H3:(g + 1) * H3, :E + Fd] = W_ih[g]
    GIw = np.full((T * B, H3), np.nan)
    _emulate(Xgm, False, Wsp, False, GIw, H3, plan.units_xproj(H3, E, 128).numpy(), 128, bias=b_ih.reshape(-1), rowmap=gm2tm)
    ref = np.stack([Xw[r] @ W_ih[gs[r % B]][:, :E].T + b_ih[gs[r % B]] for r in range(T * B)])
    assert np.allclose(GIw, ref)
    # ---- dX: A = dGI group-major (K-major, K = 3H), B = the same W array read MN-major ----
    dGI = rng.randn(T * B, H3)
    dGIgm = _gather(dGI, gm2tm, 64)
    dXw = np.full((T * B, E), np.nan)
    _emulate(dGIgm, False, Wsp, True, dXw, E, plan.units_dx(H3, E, 128).numpy(), 128, rowmap=gm2tm)
    ref = np.stack([dGI[r] @ W_ih[gs[r % B]][:, :E] for r in range(T * B)])
    assert np.allclose(dXw, ref)
    # ---- dW_ih[g] / dW_hh[g]: both operands MN-major over the group's padded K range, written into dTheta[g] ----
    XC = rng.randn(T * B, E + Fd)
    Hp = rng.randn(T * B, H)
    dGH = rng.randn(T * B, H3)
    dTheta = np.zeros((G, theta))
    _emulate(dGIgm, True, _gather(XC, gm2tm, 64), True, dTheta, E + Fd,
             plan.units_dw(H3, E + Fd, 128, theta, 0, E + Fd).numpy(), 128)
    _emulate(_gather(dGH, gm2tm, 64), True, _gather(Hp, gm2tm, 64), True, dTheta, H,
             plan.units_dw(H3, H, 128, theta, o_hh, H).numpy(), 128)
    for g in range(G):
        rows = [r for r in range(T * B) if gs[r % B] == g]
        ref_ih = dGI[rows].T @ XC[rows] if rows else np.zeros((H3, E + Fd))
        ref_hh = dGH[rows].T @ Hp[rows] if rows else np.zeros((H3, H))
        assert np.allclose(dTheta[g, :o_hh].reshape(H3, E + Fd), ref_ih)
        assert np.allclose(dTheta[g, o_hh:o_bi].reshape(H3, H), ref_hh)


def test_group_plan_cache_and_validation():
    g = torch.tensor([2, 0, 2, 1, 0])
    p1 = GroupPlan.get(g, 3, 4, "cpu")
    assert GroupPlan.get(g.clone(), 3, 4, "cpu") is p1 and GroupPlan.get(g, 3, 5, "cpu") is not p1
    assert p1.goff.tolist() == [0, 2, 3, 5] and p1.present == [0, 1, 2]
    with pytest.raises(ValueError):
        GroupPlan.get(torch.tensor([0, 3]), 3, 2, "cpu")
