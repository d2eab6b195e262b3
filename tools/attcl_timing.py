"""Debug: per-phase clock64 timestamps of attgru_cluster_fwd_kernel (needs tools/_dbg/libdbg.so built with -DCAPHN_ATTCL_TIMING)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hypernet_image_captioning_b200 import _cabi
_cabi.LIB_PATH = os.path.join(ROOT, "tools", "_dbg", "libdbg.so")
import torch
from hypernet_image_captioning_b200 import ops
B, T, Fo, E, H, P = 512, 20, 200, 200, 200, 49
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
r = lambda *s: torch.randn(*s, generator=g, device=dev)
Kp, f, GIw = r(B, P, H) * 0.5, r(B, P, Fo) * 0.5, r(T * B, 3 * H) * 0.5
Ua, W_ih, W_hh = r(H, H) / H ** 0.5, r(3 * H, E + Fo) / (E + Fo) ** 0.5, r(3 * H, H) / H ** 0.5
bu, va, bv, bhh, h0 = r(H) * 0.1, r(H) * 0.3, r(1), r(3 * H) * 0.1, r(B, H) * 0.5
Hall = torch.empty(T + 1, B, H, device=dev); Hall[0] = h0
Hbm, attn = torch.empty(B, T, H, device=dev), torch.empty(B, T, P, device=dev)
XC, saved = torch.zeros(T * B, E + Fo, device=dev), torch.empty(5, T, B, H, device=dev)
for _ in range(3):
    ops.attgru_cluster_fwd(Kp, f, GIw, Ua, bu, va, bv, W_ih, W_hh, bhh, Hall, Hbm, attn, XC, E, saved, 0, T)
torch.cuda.synchronize()
lib = _cabi.load()
buf = (ctypes.c_longlong * 32)()
lib.caphn_attcl_timestamps.argtypes = [ctypes.c_void_p]
print("rc", lib.caphn_attcl_timestamps(ctypes.cast(buf, ctypes.c_void_p)))
ts = list(buf)[:13]
all_ts = list(buf)
print("max active clusters", lib.caphn_attcl_max_clusters(H, Fo, P))
print("prologue: weights", all_ts[21]-all_ts[20], "K/f", all_ts[22]-all_ts[21], "init+sync", all_ts[23]-all_ts[22], "loop", all_ts[24]-all_ts[23])
names = ["P1 mma(h)", "u exchange", "barrier1", "scores", "softmax", "ctx+bcast", "barrier2", "ctx->bf16", "P3 mma(ctx)", "gates+bcast",
         "barrier3(+global stores)", "h->bf16"]
for i, n in enumerate(names):
    print(f"{n:28s} {ts[i + 1] - ts[i]:8d} cycles")
print("step total", ts[12] - ts[0])
