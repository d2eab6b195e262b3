"""Small driver for ncu: runs a few steps of one workload (pooled|attention, train|greedy|literal)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hypernet_image_captioning_b200 as C  # noqa: E402
from hypernet_image_captioning_b200.synth import synth_captions  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="attention")
    ap.add_argument("--mode", default="train")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--T", type=int, default=20)
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    B, T, V = a.batch, a.T, 9684
    g = torch.Generator().manual_seed(1)
    torch.manual_seed(0)
    caps = synth_captions(B, T, V, g).to(dev)
    if a.variant == "attention":
        with torch.device(dev):
            m = C.HyperNetAttention(200, 200, 200, V, None)
        feats = torch.randn(B, 49, 2048, generator=g).to(dev)

        def train():
            m.zero_grad(set_to_none=True)
            cap = m.forward(m.captioner.embed.weight[4:5])
            loss, _, _ = cap.forward_loss(feats, caps, 0.0, ignore_index=0)
            loss.backward()

        def greedy():
            with torch.no_grad():
                m.forward(m.captioner.embed.weight[4:5])(feats, caps, 1.0)
    else:
        with torch.device(dev):
            m = C.HyperNetPooled(200, 150, V, None)
        pooled = torch.relu(torch.randn(B, 2048, generator=g)).to(dev)
        h0 = torch.rand(B, 150, generator=g).to(dev)

        def train():
            m.zero_grad(set_to_none=True)
            cap = m.forward(m.captioner.embed.weight[4:5])
            loss, _ = cap.forward_loss(m.image_encoder(pooled), caps, h0=h0)
            loss.backward()

        def greedy():
            with torch.no_grad():
                m.forward(m.captioner.embed.weight[4:5]).infer(m.image_encoder(pooled), max_len=T, h0=h0)
    if a.mode == "literal":
        m.grad_mode = "literal"
    m.async_hypernet = os.environ.get("CAPHN_ASYNC_HN", "1") != "0"
    fn = greedy if a.mode == "greedy" else train
    for _ in range(a.warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{a.variant} {a.mode}: {e0.elapsed_time(e1) / a.steps:.3f} ms/step, {B * a.steps / e0.elapsed_time(e1) * 1e3:.0f} captions/s")


if __name__ == "__main__":
    main()
