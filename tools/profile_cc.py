"""One eager many-style training step (configs[3], G domains in the batch) for an ncu launch list: `python tools/profile_cc.py [G]`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hypernet_image_captioning_b200 as C  # noqa: E402
from hypernet_image_captioning_b200.synth import synth_captions  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 100
he = max(G, 100)
B, T, V = 512, 20, 9684
dev = torch.device("cuda", 0)
torch.manual_seed(0)
with torch.device(dev):
    m = C.HyperNetAttention(200, 200, 200, V, None, cc=True, hyper_emb=he)
g = torch.Generator().manual_seed(1)
feats = torch.randn(B, 49, 2048, generator=g).to(dev)
caps = synth_captions(B, T, V, g).to(dev)
groups = torch.arange(B) % G
styles = torch.eye(he, device=dev)[:G].contiguous()


def step():
    m.zero_grad(set_to_none=True)
    cap = m.forward_grouped(styles)
    loss, _, _ = cap.forward_loss(feats, caps, 0.0, ignore_index=0, groups=groups)
    loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
torch.cuda.synchronize()
print(f"G={G}: {e0.elapsed_time(e1):.3f} ms eager")
