#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel and per (kernel, grid) totals.
    python tools/launch_summary.py gpurun_out/launches.csv [first_id last_id]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    by_grid = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
            i = int(row["ID"])
        except Exception:
            continue
        if not (lo <= i < hi):
            continue
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v      # -> us
        short = re.sub(r"\(.*", "", row["Kernel Name"])
        short = re.sub(r"^void ", "", short)[:64]
        agg[short][0] += 1
        agg[short][1] += v
        by_grid[(short, row["Grid Size"])][0] += 1
        by_grid[(short, row["Grid Size"])][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f %% |" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
    print("\nTotal %.2f ms over %d launches\n" % (tot / 1e3, n))
    print("| kernel | grid | launches | avg us |\n|---|---|---:|---:|")
    for (k, g), v in sorted(by_grid.items(), key=lambda kv: -kv[1][1])[:30]:
        print("| `%s` | %s | %d | %.1f |" % (k, g, v[0], v[1] / v[0]))


if __name__ == "__main__":
    main()
