"""configs[3] lines only (the cc_extras section of bench.py) on one GPU: `python tools/bench_cc.py`."""
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def timed(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


if __name__ == "__main__":
    args = types.SimpleNamespace(batch=512, steps=20)
    rf = {}
    peak = 6537.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    out = bench.cc_extras(args, torch.device("cuda", 0), 1, timed, rf, peak)
    for k, v in out.items():
        print(f"{k}: {v:.0f}")
    for k, v in rf.items():
        print(f"{k}: {v['ms']:.3f} ms, {v['frac']:.3f} of the HBM roofline (floor {v['floor_ms']:.3f} ms)")
