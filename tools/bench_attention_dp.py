"""configs[2] (attention variant) under data parallelism: the attention section of bench.py's extras on N ranks (eager and
graphed training step with the d(theta) all-reduce + overlapped shared-gradient bucket, greedy decode), timed like bench.py
(barrier + synchronize on both sides, max over ranks).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/bench_attention_dp.py
"""
import json
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    args = types.SimpleNamespace(batch=512, steps=20)
    peak = 6537.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    rf = {}
    out = bench.attention_extras(args, dev, world, timed, rf, peak)
    if rank == 0:
        print(f"attention variant, {world} x B200, B=512/GPU, T=20 (whole-job captions/s)")
        for k, v in out.items():
            print(f"  {k}: {v:.0f}")
        for k, v in rf.items():
            print(f"  {k}: {v['ms']:.3f} ms/step")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
