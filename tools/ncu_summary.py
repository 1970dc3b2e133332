#!/usr/bin/env python
"""Summarise `ncu --set full` raw pages (ncu -i x.ncu-rep --page raw --csv) into a markdown table and a
per-kernel DRAM-traffic JSON that bench.py reads for roofline.traffic.
    python tools/ncu_summary.py out.md traffic.json raw1.csv [raw2.csv ...]"""
import csv
import json
import re
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem)"),
]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def main():
    out_md, out_json, paths = sys.argv[1], sys.argv[2], sys.argv[3:]
    lines = ["| capture | # | kernel | grid | " + " | ".join(c[1] for c in COLS) + " |",
             "|---|---:|---|---|" + "---:|" * len(COLS)]
    traffic = {}
    for path in paths:
        rows = [r for r in csv.reader(open(path)) if r]
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
            cells = []
            num = {}
            for key, _ in COLS:
                if key not in ix:
                    cells.append("-")
                    continue
                v, u = r[ix[key]], units[ix[key]]
                try:
                    f = float(v.replace(",", ""))
                except ValueError:
                    cells.append(v)
                    continue
                if u in SCALE:
                    f *= SCALE[u]
                    num[key] = f
                    cells.append("%.1f us" % f if "time" in key else "%.1f MB" % (f / 1e6))
                else:
                    cells.append("%.1f" % f)
            lines.append("| %s | %s | `%s` | %s | %s |" % (path.split("/")[-1], r[ix["ID"]], name, r[ix["Grid Size"]],
                                                         " | ".join(cells)))
            t = traffic.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
            t["launches"] += 1
            t["dram_bytes"] += num.get("dram__bytes_read.sum", 0) + num.get("dram__bytes_write.sum", 0)
            t["us"] += num.get("gpu__time_duration.sum", 0)
    open(out_md, "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(out_json, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
