"""Pieces of one pooled greedy-decode step (B = 512, H = 150, V = 9684), each as 20 dependent launches replayed from a CUDA
graph: the fused GRU step, the vocabulary GEMM with arg-max partials, both alternating.  `python tools/bench_decode_parts.py`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernet_image_captioning_b200 import ops  # noqa: E402

B, H, V, T = 512, 150, 9684, 20
g = torch.Generator().manual_seed(0)
table = torch.randn(V, 3 * H, generator=g).cuda()
W_hh = (torch.randn(3 * H, H, generator=g) * 0.1).cuda()
b_hh = torch.randn(3 * H, generator=g).cuda()
fc_w = (torch.randn(V, H, generator=g) * 0.2).cuda()
fc_b = torch.randn(V, generator=g).cuda()
h = [torch.rand(B, H, generator=g).cuda(), torch.empty(B, H, device="cuda")]
nslot = 2 * ((V + 127) // 128)
pv = torch.randn(B, nslot, generator=g).cuda()
pi = torch.randint(0, V, (B, nslot), generator=g, dtype=torch.int32).cuda()
Kp = ops.round64(H)
hi = torch.zeros(B, Kp, device="cuda", dtype=torch.bfloat16)
lo = torch.zeros(B, Kp, device="cuda", dtype=torch.bfloat16)
hop = ops.SplitOperand(hi, lo, B, H, Kp)
wop = ops.split_bf16(fc_w)
out = torch.empty(B, T, V, device="cuda")
state = {"np": nslot - 1}


def step(t):
    ops.gru_decode_step(None, pv, pi, state["np"], table, W_hh, b_hh, h[t & 1], h[(t + 1) & 1], hi, lo)


def gemm(t):
    state["np"] = ops.gemm_tc_amax(hop, wop, fc_b, out[:, t, :], pv, pi)


def timed(fn, reps=10):
    for t in range(2):
        fn(t)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(0)
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for t in range(T):
            fn(t)
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (T * reps) * 1e3


gemm(0)
print(f"fused GRU step alone:          {timed(step):6.1f} us per launch (20 dependent launches, graph replay)")
print(f"vocabulary GEMM + arg-max:     {timed(gemm):6.1f} us")
print(f"step + GEMM alternating:       {timed(lambda t: (step(t), gemm(t))):6.1f} us per pair")
for bn in (128, 144, 160, 192, 256):
    os.environ["CAPHN_TC_BN_FORCE"] = str(bn)
    print(f"  GEMM with BN={bn}: {timed(gemm):6.1f} us")
os.environ.pop("CAPHN_TC_BN_FORCE", None)
os.environ["CAPHN_TC_TMA_STORE"] = "0"
print(f"  GEMM, default epilogue (no TMA store): {timed(gemm):6.1f} us")
os.environ.pop("CAPHN_TC_TMA_STORE", None)
