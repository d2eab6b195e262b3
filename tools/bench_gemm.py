"""N-tile sweep of the tcgen05 GEMM on the vocabulary-projection shapes: `python tools/bench_gemm.py`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernet_image_captioning_b200 import ops  # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


g = torch.Generator().manual_seed(0)
for (M, N, K) in [(10240, 9684, 150), (10240, 9684, 200), (512, 9684, 150)]:
    X = torch.randn(M, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    b = torch.randn(N, generator=g).cuda()
    xs, ws = ops.split_bf16(X), ops.split_bf16(W)
    out = torch.empty(M, N, device="cuda")
    ref = None
    for bn in (0, 128, 144, 160, 192, 224, 256):
        if bn:
            os.environ["CAPHN_TC_BN_FORCE"] = str(bn)
        else:
            os.environ.pop("CAPHN_TC_BN_FORCE", None)
        us = timed(lambda: ops.gemm_tc(xs, ws, bias=b, out=out))
        if ref is None:
            ref = out.clone()
        err = (out - ref).abs().max().item()
        print(f"M={M} N={N} K={K} BN={'auto' if not bn else bn}: {us:8.1f} us  ({4.0 * M * N / us / 1e3:6.0f} GB/s of output)  "
              f"max|diff vs auto| {err:.1e}", flush=True)
os.environ.pop("CAPHN_TC_BN_FORCE", None)

# TMA-store epilogue (CAPHN_TC_TMA_STORE=1; N % 4 == 0 only) against the default epilogue, same shapes
for (M, N, K) in [(10240, 9684, 150), (10240, 9684, 200), (512, 9684, 200), (25088, 200, 2048)]:
    X = torch.randn(M, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    b = torch.randn(N, generator=g).cuda()
    xs, ws = ops.split_bf16(X), ops.split_bf16(W)
    outs = []
    for tma in ("0", "1"):
        os.environ["CAPHN_TC_TMA_STORE"] = tma
        out = torch.empty(M, N, device="cuda")
        us = timed(lambda: ops.gemm_tc(xs, ws, bias=b, out=out))
        outs.append(out)
        print(f"M={M} N={N} K={K} tma_store={tma}: {us:8.1f} us  ({4.0 * M * N / us / 1e3:6.0f} GB/s of output)", flush=True)
    print(f"    max|tma - default| = {(outs[0] - outs[1]).abs().max().item():.1e}")
os.environ.pop("CAPHN_TC_TMA_STORE", None)
