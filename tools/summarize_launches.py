"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: last N launches aggregated by kernel."""
import collections
import csv
import re
import sys


def main(path, last=None, top=30):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    if last:
        rows = rows[-int(last):]
    tot = sum(float(r["Metric Value"]) for r in rows) / 1e3
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("caphn::", "")
        agg[k][0] += 1
        agg[k][1] += float(r["Metric Value"]) / 1e3
    print(f"{len(rows)} launches, {tot:.1f} us")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(top)]:
        print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  x{v[0]:<4d} {k[:90]}")


if __name__ == "__main__":
    main(*sys.argv[1:])
