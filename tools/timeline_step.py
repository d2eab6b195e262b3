"""In-graph kernel timeline of the pooled train step (B=512, T=20): the step is captured as ONE CUDA graph, replayed under
torch.profiler (CUPTI records the kernels of a graph launch), and the kernels of one replay are listed in start order with
their stream, duration and the gap to the previous kernel END on any stream -- shows what the ncu launch list cannot: which
launches overlap and where the device idles.   python tools/timeline_step.py [pooled|attention]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hypernet_image_captioning_b200 as C  # noqa: E402
from hypernet_image_captioning_b200 import graphs  # noqa: E402
from hypernet_image_captioning_b200.synth import synth_captions  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "pooled"
dev = torch.device("cuda", 0)
B, T, V = 512, 20, 9684
g = torch.Generator().manual_seed(1)
torch.manual_seed(0)
caps = synth_captions(B, T, V, g).to(dev)
if variant == "attention":
    with torch.device(dev):
        m = C.HyperNetAttention(200, 200, 200, V, None)
    feats = torch.randn(B, 49, 2048, generator=g).to(dev)
    m.async_hypernet = True

    def step():
        m.zero_grad(set_to_none=True)
        cap = m.forward(m.captioner.embed.weight[4:5])
        loss, _, _ = cap.forward_loss(feats, caps, 0.0, ignore_index=0)
        loss.backward()
        return loss
else:
    with torch.device(dev):
        m = C.HyperNetPooled(200, 150, V, None)
    pooled = torch.relu(torch.randn(B, 2048, generator=g)).to(dev)
    h0 = torch.rand(B, 150, generator=g).to(dev)
    m.async_hypernet = os.environ.get("CAPHN_ASYNC_HN", "1") != "0"     # as bench.py runs the headline

    def step():
        m.zero_grad(set_to_none=True)
        cap = m.forward(m.captioner.embed.weight[4:5])
        loss, _ = cap.forward_loss(m.image_encoder(pooled), caps, h0=h0)
        loss.backward()
        return loss

gs = graphs.GraphedStep(step, (), params=list(m.parameters()), release=m.release_graph)
assert gs.captured
for _ in range(3):
    gs()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        gs()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
# split into replays by the largest gaps
starts = [e.time_range.start for e in ev]
n = len(ev) // 3
one = ev[n:2 * n]
t0 = one[0].time_range.start
end_prev = t0
busy = 0.0
print(f"# {variant} train step, one graph replay: {len(one)} device activities, "
      f"{(one[-1].time_range.end - t0):.1f} us from first start to last end")
print(f"# {'start':>8s} {'dur':>8s} {'gap':>7s}  stream  kernel")
streams = {}
for e in one:
    s = streams.setdefault(getattr(e, 'device_index', 0) * 1000 + (e.device_resource_id if hasattr(e, 'device_resource_id') else 0), len(streams))
    st, en = e.time_range.start - t0, e.time_range.end - t0
    gap = e.time_range.start - end_prev
    name = e.name.replace("void ", "").replace("caphn::", "")[:70]
    print(f"  {st:8.1f} {en - st:8.1f} {gap:7.1f}  s{s:<5d} {name}")
    end_prev = max(end_prev, e.time_range.end)
