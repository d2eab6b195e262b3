"""Summarise `ncu --set full` reports (read with `ncu -i X.ncu-rep --page raw --csv`): one line per profiled launch with
duration, DRAM bytes / throughput, tensor-pipe activity, occupancy, registers.

    python tools/summarize_ncu.py gpurun_out/prof_gemm.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_full_summary.txt
"""
import csv
import io
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "us", 1e-3),
    ("dram__bytes_read.sum", "rdMB", 1e-6),
    ("dram__bytes_write.sum", "wrMB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1.0),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1.0),
    ("TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "hmma_cyc", 1.0),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tmem%", 1.0),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1.0),
    ("lts__t_sector_hit_rate.pct", "L2hit%", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
    ("launch__grid_size", "grid", 1.0),
]
UNIT_SCALE = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9,
              "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def main(paths):
    for path in paths:
        hdr, units, data = load(path)
        idx = {h: i for i, h in enumerate(hdr)}
        tens = sorted(h for h in hdr if "tensor" in h)
        print(f"# {path}: {len(data)} profiled launches (ncu --set full --clock-control none; one replay set per launch)")
        print("# tensor% = sm__pipe_tensor_cycles_active (pct of peak, active cycles); hmma_cyc = tensor-pipe HMMA-subpipe active cycles"
              " (realtime counter, per TPC = 2 SMs: UTCHMMA / HMMA work shows up here); tmem% = sm__mem_tensor_cycles_active")
        cols = [w for w in WANT if w[0] in idx]
        print(f"{'kernel':44s} " + " ".join(f"{c[1]:>10s}" for c in cols))
        for r in data:
            name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("caphn::", "")
            vals = []
            for m, _, sc in cols:
                v = r[idx[m]].replace(",", "")
                try:
                    x = float(v) * UNIT_SCALE.get(units[idx[m]], 1.0) * sc
                    vals.append(f"{x:10.1f}")
                except ValueError:
                    vals.append(f"{v[:10]:>10s}")
            print(f"{name[:44]:44s} " + " ".join(vals))
        print()


if __name__ == "__main__":
    main(sys.argv[1:])
