"""Debug: per-phase clock64 timestamps of the step-split kernels (needs tools/_dbg/libdbg.so built with -DCAPHN_ATTCL_TIMING)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hypernet_image_captioning_b200 import _cabi
_cabi.LIB_PATH = os.path.join(ROOT, "tools", "_dbg", "libdbg.so")
import torch
from hypernet_image_captioning_b200 import ops
B, T, Fo, E, H, P = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 20, 200, 200, 200, 49
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
r = lambda *s: torch.randn(*s, generator=g, device=dev)
Kp, f, GIw = r(B, P, H) * 0.5, r(B, P, Fo) * 0.5, r(T * B, 3 * H) * 0.5
Ua, W_ih, W_hh = r(H, H) / H ** 0.5, r(3 * H, E + Fo) / (E + Fo) ** 0.5, r(3 * H, H) / H ** 0.5
bu, va, bv, bhh, h0 = r(H) * 0.1, r(H) * 0.3, r(1), r(3 * H) * 0.1, r(B, H) * 0.5
Hall = torch.empty(T + 1, B, H, device=dev); Hall[0] = h0
Hbm, attn = torch.empty(B, T, H, device=dev), torch.empty(B, T, P, device=dev)
XC, saved = torch.zeros(T * B, E + Fo, device=dev), torch.empty(5, T, B, H, device=dev)
lw = ops.AttGruWeights(W_ih, W_hh, Ua, E, P, step=True)
for _ in range(3):
    ops.attgru_fwd(Kp, f, GIw, lw, bu, va, bv, bhh, Hall, Hbm, attn, XC, E, saved, 0, T)
torch.cuda.synchronize()
lib = _cabi.load()
buf = (ctypes.c_longlong * 32)()
lib.caphn_attst_timestamps.argtypes = [ctypes.c_void_p]
print("rc", lib.caphn_attst_timestamps(ctypes.cast(buf, ctypes.c_void_p)))
ts = list(buf)
an = ["prologue + dependency wait", "load u", "wait K/f tiles", "scores", "softmax", "ctx"]
for i, n in enumerate(an):
    print(f"A {n:28s} {ts[i + 1] - ts[i]:8d} cycles")
print("A total", ts[6] - ts[0])
yn = ["prefetch + dependency wait", "operand copies", "mma", "gates"]
for i, n in enumerate(yn):
    print(f"Y {n:28s} {ts[17 + i] - ts[16 + i]:8d} cycles")
print("Y total", ts[20] - ts[16])
