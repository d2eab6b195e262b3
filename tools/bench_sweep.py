"""BASELINE configs[4]: batch / sequence / hidden-size sweep of the attention variant (hypernet_attention.HyperNet +
AttentionGru), train (teacher-forced fwd+bwd, flow mode) and greedy decode, one GPU.  Prints a table with captions/s and
the fraction of the compulsory-traffic HBM roofline.

    python tools/bench_sweep.py [--quick] > profiles/rNN_sweep.txt

E = F = 200 (the launchers' defaults) unless H >= 512 rows say otherwise; V = 9684, P = 49, D = 2048.  Heads: H = 200 ->
0.58 GB, H = 512 -> 7.97 GB, H = 1024 -> 91 GB (fp32) -- the last one is run with bf16 heads (hypernet.set_precision).
Shapes with H > 208 run the L2-streaming recurrence (the step-split kernels keep the weight fragments in registers)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypernet_image_captioning_b200 as C  # noqa: E402
from hypernet_image_captioning_b200 import graphs, ops  # noqa: E402
from hypernet_image_captioning_b200.synth import synth_captions  # noqa: E402

V, P, D = 9684, 49, 2048


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run_point(B, T, H, E=200, Fd=200, bf16=False, peak=6537.6, steps=10):
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with torch.device(dev):
        m = C.HyperNetAttention(Fd, E, H, V, None)
    if bf16:
        m.set_precision("bf16")
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(B, P, D, generator=g).to(dev)
    caps = synth_captions(B, T, V, g).to(dev)
    style = lambda: m.captioner.embed.weight[4:5]

    def train():
        m.zero_grad(set_to_none=True)
        loss, _, _ = m.forward(style()).forward_loss(feats, caps, 0.0, ignore_index=0)
        loss.backward()

    def greedy():
        with torch.no_grad():
            return m.forward(style())(feats, caps, 1.0)

    run = train
    m.async_hypernet = True
    gtrain = graphs.GraphedStep(train, (), params=list(m.parameters()), release=m.release_graph)
    m.async_hypernet = False
    if gtrain.captured:
        run = gtrain
    ms_tr = timed(run, steps)
    run = gtrain = None
    m.zero_grad(set_to_none=True)
    torch.cuda.empty_cache()
    ms_de = timed(greedy, steps)
    s = 2.0 if bf16 else 4.0
    hn_b = sum(p.numel() * p.element_size() for n, p in m.named_parameters() if n.startswith("hn_"))
    w2_b = sum(h[2].weight.numel() * h[2].weight.element_size() for h in m.hn_heads)
    sh_b = 4.0 * sum(p.numel() for n, p in m.named_parameters() if not n.startswith("hn_") and not n.startswith("captioner.gru"))
    per_cap = P * D * 4.0 + 8.0 * T + T * V * 4.0 + T * P * 4.0
    tr_b = 2 * hn_b + w2_b + 2 * sh_b + B * (per_cap + T * V * 4.0)
    de_b = hn_b + sh_b + B * per_cap
    graphs.clear()
    ops.set_precision("fp32")
    del m
    torch.cuda.empty_cache()
    return {"B": B, "T": T, "H": H, "E": E, "F": Fd, "heads": "bf16" if bf16 else "fp32", "head_GB": hn_b / 1e9,
            "train_ms": ms_tr, "train_cps": B / ms_tr * 1e3, "train_frac": tr_b / (peak * 1e9) / (ms_tr * 1e-3),
            "decode_ms": ms_de, "decode_cps": B / ms_de * 1e3, "decode_frac": de_b / (peak * 1e9) / (ms_de * 1e-3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    peak = 6537.6
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                 "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    pts = [(64, 20, 200), (512, 16, 200), (512, 32, 200), (512, 64, 200), (2048, 20, 200), (4096, 20, 200),
           (512, 20, 512), (512, 64, 512), (2048, 20, 512)]
    if a.quick:
        pts = [(64, 20, 200), (512, 64, 200), (512, 20, 512)]
    print(f"# attention variant sweep, 1 x B200, fp32 unless noted, HBM peak {peak:.0f} GB/s; frac = compulsory bytes / peak / time")
    print(f"{'B':>5} {'T':>3} {'H':>5} {'heads':>6} {'GB':>6} | {'train ms':>9} {'capt/s':>9} {'frac':>5} | {'decode ms':>9} {'capt/s':>9} {'frac':>5}")
    rows = []
    for B, T, H in pts:
        for bf in ((False,) if H < 1024 else (True,)):
            try:
                r = run_point(B, T, H, bf16=bf, peak=peak)
            except Exception as e:  # noqa: BLE001
                print(f"{B:5d} {T:3d} {H:5d}  FAILED: {type(e).__name__}: {str(e)[:120]}")
                torch.cuda.empty_cache()
                continue
            rows.append(r)
            print(f"{B:5d} {T:3d} {H:5d} {r['heads']:>6} {r['head_GB']:6.2f} | {r['train_ms']:9.3f} {r['train_cps']:9.0f} "
                  f"{r['train_frac']:5.2f} | {r['decode_ms']:9.3f} {r['decode_cps']:9.0f} {r['decode_frac']:5.2f}", flush=True)
    # one bf16-head point at H = 512 for comparison, and H = 1024 (91 GB of fp32 heads: bf16 storage only)
    for B, T, H in ([(512, 20, 512)] if a.quick else [(512, 20, 512), (512, 20, 1024)]):
        try:
            r = run_point(B, T, H, bf16=True, peak=peak, steps=5)
            rows.append(r)
            print(f"{B:5d} {T:3d} {H:5d} {r['heads']:>6} {r['head_GB']:6.2f} | {r['train_ms']:9.3f} {r['train_cps']:9.0f} "
                  f"{r['train_frac']:5.2f} | {r['decode_ms']:9.3f} {r['decode_cps']:9.0f} {r['decode_frac']:5.2f}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{B:5d} {T:3d} {H:5d}  bf16 FAILED: {type(e).__name__}: {str(e)[:160]}")
            torch.cuda.empty_cache()
    print("# json: " + json.dumps(rows))


if __name__ == "__main__":
    main()
