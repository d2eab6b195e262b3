"""tcgen05 GEMM (bf16x3 split, fp32 accumulate in TMEM) against fp64 matmul."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from golden_util import rel_err


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 150, 150), (1000, 9684, 150), (2000, 150, 9684),
                                   (9684, 150, 2048), (300, 200, 2048), (128, 128, 6400), (4097, 257, 65)])
def test_gemm_tc_split_matches_fp64(M, N, K):
    from hypernet_image_captioning_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    b = torch.randn(N, generator=g).cuda()
    out = ops.gemm_tc(ops.split_bf16(A), ops.split_bf16(W), bias=b)
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t() + b.double()
    # The tensor core adds each MMA into the fp32 TMEM accumulator with truncation, so the error grows with the number
    # of accumulation steps (K/16 * 3): ~1e-6 at K = 150-200 (logits), ~5e-6 at K = 2048, ~4e-5 at K ~ 1e4 (the two
    # vocabulary-projection gradient products, whose budget is 1e-3).
    tol = 2e-5 if K <= 2048 else 1e-4
    assert rel_err(out, ref) < tol
    out2 = ops.gemm_tc(ops.split_bf16(A), ops.split_bf16(W), bias=b, relu=True)
    assert rel_err(out2, ref.relu()) < tol


def test_gemm_tc_transposed_operands_and_bf16_mode():
    from hypernet_image_captioning_b200 import ops
    g = torch.Generator().manual_seed(5)
    dl = torch.randn(1536, 700, generator=g).cuda()      # [K, M]
    Hs = torch.randn(1536, 150, generator=g).cuda()      # [K, N]
    out = ops.gemm_tc(ops.split_bf16_t(dl), ops.split_bf16_t(Hs))   # dl^T @ Hs
    assert rel_err(out, dl.double().t() @ Hs.double()) < 2e-5  # K = 1536
    A = torch.randn(512, 256, generator=g).cuda()
    W = torch.randn(384, 256, generator=g).cuda()
    out = ops.gemm_tc(ops.split_bf16(A, want_lo=False), ops.split_bf16(W, want_lo=False))
    assert rel_err(out, A.double() @ W.double().t()) < 2e-2


def test_gemm_tc_strided_output():
    from hypernet_image_captioning_b200 import ops
    g = torch.Generator().manual_seed(6)
    A = torch.randn(256, 100, generator=g).cuda()
    W = torch.randn(200, 100, generator=g).cuda()
    big = torch.zeros(256, 3, 200).cuda()
    ops.gemm_tc(ops.split_bf16(A), ops.split_bf16(W), out=big[:, 1, :])
    assert rel_err(big[:, 1, :], A.double() @ W.double().t()) < 2e-5
    assert float(big[:, 0, :].abs().max()) == 0 and float(big[:, 2, :].abs().max()) == 0


@pytest.mark.parametrize("M,N,K,splitk", [(512, 150, 9684, 0), (450, 150, 10240, 0), (450, 200, 4096, 7), (9684, 150, 4096, 0),
                                          (256, 208, 640, 3), (300, 16, 128, 1), (130, 250, 200, 0), (1024, 600, 200, 0)])
def test_gemm_tc_tile_shapes_and_splitk(M, N, K, splitk):
    """Run-time N tile (N <= 256 -> one tile rounded to 16) and split-K with atomic accumulation."""
    from hypernet_image_captioning_b200 import ops
    g = torch.Generator().manual_seed(M * 3 + N * 5 + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    b = torch.randn(N, generator=g).cuda()
    out = ops.gemm_tc(ops.split_bf16(A), ops.split_bf16(W), bias=b, splitk=splitk)
    ref = A.double() @ W.double().t() + b.double()
    assert rel_err(out, ref) < (2e-5 if K <= 2048 else 1e-4)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (700, 150, 1536), (9684, 150, 2048), (200, 2048, 3000), (257, 130, 100)])
def test_gemm_tc_mn_major_operands(M, N, K):
    """Transposed products read in place: A and/or B stored [K, MN] (MN-major UMMA descriptors, 64x64 TMA boxes)."""
    from hypernet_image_captioning_b200 import ops
    g = torch.Generator().manual_seed(M + 2 * N + 3 * K)
    At = torch.randn(K, M, generator=g).cuda()      # A^T  (A is [M, K])
    Bt = torch.randn(K, N, generator=g).cuda()      # B^T  (B is [N, K])
    ref = At.double().t() @ Bt.double()
    tol = 2e-5 if K <= 2048 else 1e-4
    # A MN-major, B K-major (via transposing split)
    out = ops.gemm_tc(ops.split_bf16(At, mn=True), ops.split_bf16_t(Bt))
    assert rel_err(out, ref) < tol
    # A K-major, B MN-major
    out = ops.gemm_tc(ops.split_bf16_t(At), ops.split_bf16(Bt, mn=True))
    assert rel_err(out, ref) < tol
    # both MN-major
    out = ops.gemm_tc(ops.split_bf16(At, mn=True), ops.split_bf16(Bt, mn=True))
    assert rel_err(out, ref) < tol


@pytest.mark.parametrize("M,N,K,relu", [(10240, 9684, 150, False), (512, 9684, 150, False), (1000, 450, 200, False),
                                        (300, 200, 2048, True), (4097, 257, 65, False), (128, 160, 64, True)])
def test_gemm_tc_tma_store_epilogue_is_bit_identical(M, N, K, relu, monkeypatch):
    """The TMA-store epilogue (default for N % 4 == 0, no split-K) must write exactly what the register-store epilogue
    (CAPHN_TC_TMA_STORE=0) writes, incl. bias, ReLU and N tiles that are not a multiple of 32 columns, and nothing outside
    [M, N]; ragged-N shapes take the register-store epilogue under both settings."""
    from hypernet_image_captioning_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    b = torch.randn(N, generator=g).cuda()
    xs, ws = ops.split_bf16(A), ops.split_bf16(W)
    ldc = ((N + 3) // 4) * 4 + 8
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("CAPHN_TC_TMA_STORE", flag)
        buf = torch.full((M + 2, ldc), -7.0, device="cuda")
        ops.gemm_tc(xs, ws, bias=b, relu=relu, out=buf[1:M + 1, :N])
        torch.cuda.synchronize()
        outs.append(buf)
    assert torch.equal(outs[0], outs[1])
    assert float(outs[1][0].max()) == -7.0 and float(outs[1][-1].max()) == -7.0 and float(outs[1][:, N:].max()) == -7.0
