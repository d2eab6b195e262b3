"""Single GRU step (decode) on the two recurrence kernels: L2-streaming gru_seq_fwd vs weights-resident gru_cluster_fwd, T = 1."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernet_image_captioning_b200 import ops  # noqa: E402

B, H = 512, 150
g = torch.Generator().manual_seed(0)
GI = torch.randn(B, 3 * H, generator=g).cuda()
W_hh = (torch.randn(3 * H, H, generator=g) * 0.1).cuda()
b_hh = torch.randn(3 * H, generator=g).cuda()
h = torch.rand(B, H, generator=g).cuda()
WhhT = ops.transpose_pad(W_hh, ops.round4(3 * H))


def timed(fn, n=20, reps=20):
    """n dependent calls captured into one CUDA graph (device time without host launch overhead), replayed reps times."""
    for _ in range(3):
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


a = ops.gru_seq_fwd(GI, WhhT, b_hh, h, 1, save=False, want_bm=False)[0][1]
b = ops.gru_cluster_fwd(GI, W_hh, b_hh, h, 1, save=False, want_bm=False)[0][1]
print("max diff", (a - b).abs().max().item())
print(f"gru_seq_fwd T=1: {timed(lambda: ops.gru_seq_fwd(GI, WhhT, b_hh, h, 1, save=False, want_bm=False)):.1f} us (graph replay)")
print(f"gru_cluster_fwd T=1: {timed(lambda: ops.gru_cluster_fwd(GI, W_hh, b_hh, h, 1, save=False, want_bm=False)):.1f} us")
