#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`
launch list per kernel name.

    python tools/ncu_summary.py gpurun_out/launches.csv [--step K] > profiles/rNN_launches.md

--step K keeps only the K-th training step (0-based) of the list: a step starts at the first of the two
im2col_kernel launches (video stem, audio stem) and runs to the next step's first im2col_kernel.
--last N / --first N keep the last / first N launches instead."""
import argparse
import csv
import io
import json
import re
from collections import defaultdict

UNIT_NS = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}
UNIT_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(path):
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    launches = {}
    order = []
    for r in csv.DictReader(io.StringIO(text[start:])):
        i = int(r["ID"])
        if i not in launches:
            launches[i] = {"name": re.sub(r"\(.*", "", r["Kernel Name"]).strip().replace("<unnamed>::", "").replace("void ", ""), "ns": 0.0, "rd": None, "wr": None}
            order.append(i)
        val = float(r["Metric Value"].replace(",", ""))
        m, u = r["Metric Name"], r.get("Metric Unit", "")
        if m.startswith("gpu__time_duration"):
            launches[i]["ns"] = val * UNIT_NS.get(u, 1.0)
        elif m.startswith("dram__bytes_read"):
            launches[i]["rd"] = val * UNIT_B.get(u, 1.0)
        elif m.startswith("dram__bytes_write"):
            launches[i]["wr"] = val * UNIT_B.get(u, 1.0)
    return [launches[i] for i in order]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--last", type=int, default=0)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--step", type=int, default=None)
    ap.add_argument("--json", default=None, help="also write {kernel: {launches, ms, dram_bytes}} here")
    args = ap.parse_args()
    rows = load(args.csv)
    if args.step is not None:
        marks = [i for i, r in enumerate(rows) if r["name"].startswith("im2col_kernel")][::2]
        lo = marks[args.step]
        hi = marks[args.step + 1] if args.step + 1 < len(marks) else len(rows)
        rows = rows[lo:hi]
    if args.first:
        rows = rows[: args.first]
    if args.last:
        rows = rows[-args.last:]
    agg = defaultdict(lambda: [0, 0.0, 0.0, False])
    total = 0.0
    for r in rows:
        a = agg[r["name"]]
        a[0] += 1
        a[1] += r["ns"]
        if r["rd"] is not None:
            a[2] += r["rd"] + (r["wr"] or 0.0)
            a[3] = True
        total += r["ns"]
    has_dram = any(a[3] for a in agg.values())
    print(f"launches: {len(rows)}   total device time: {total / 1e6:.3f} ms (ncu: serialised, cold cache — compare shares)\n")
    print("| kernel | launches | total ms | share | avg us |" + (" DRAM MB (rd+wr) | DRAM GB/s |" if has_dram else ""))
    print("|---|---:|---:|---:|---:|" + ("---:|---:|" if has_dram else ""))
    for name, (n, ns, by, _) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = f"| `{name[:90]}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.1f} |"
        if has_dram:
            line += f" {by / 1e6:.1f} | {by / ns:.0f} |" if ns else " | |"
        print(line)
    if args.json:
        with open(args.json, "w") as f:
            json.dump({k: {"launches": v[0], "ms": v[1] / 1e6, "dram_bytes": v[2]} for k, v in agg.items()}, f, indent=1)


if __name__ == "__main__":
    main()
