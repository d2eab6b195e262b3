"""GRU recurrence kernels at the benchmark shape (B = 512, T = 20, H = 150): L2-streaming / cluster-resident / CTA-resident,
forward and BPTT, graph-replayed.  `python tools/bench_gru.py`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernet_image_captioning_b200 import ops  # noqa: E402

B, T, H = 512, 20, 150
g = torch.Generator().manual_seed(0)
GI = torch.randn(T * B, 3 * H, generator=g).cuda()
W = (torch.randn(3 * H, H, generator=g) * 0.1).cuda()
b = torch.randn(3 * H, generator=g).cuda()
h0 = torch.rand(B, H, generator=g).cuda()
dH = torch.randn(B, T, H, generator=g).cuda()
WT = ops.transpose_pad(W, ops.round4(3 * H))
Wp = ops.copy_pad(W, ops.round4(H))


def timed(fn, n=5, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3


Hall, Hbm, sv, _ = ops.gru_seq_fwd(GI, WT, b, h0, T)
print(f"forward : streaming {timed(lambda: ops.gru_seq_fwd(GI, WT, b, h0, T)):7.1f} us | cluster "
      f"{timed(lambda: ops.gru_cluster_fwd(GI, W, b, h0, T)):7.1f} us | CTA-resident "
      f"{timed(lambda: ops.gru_resident_fwd(GI, W, b, h0, T)):7.1f} us")
print(f"backward: streaming {timed(lambda: ops.gru_seq_bwd(dH, sv, Hall, None, Wp)):7.1f} us | cluster "
      f"{timed(lambda: ops.gru_cluster_bwd(dH, sv, Hall, W)):7.1f} us | CTA-resident "
      f"{timed(lambda: ops.gru_resident_bwd(dH, sv, Hall, W)):7.1f} us")
