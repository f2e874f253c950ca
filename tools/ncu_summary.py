#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.

    python tools/ncu_summary.py gpurun_out/launches.csv [--last N] > profiles/rNN_launches.md

--last N keeps only the last N launches (e.g. one training step) before aggregating."""
import argparse
import csv
import io
import re
import sys
from collections import defaultdict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--last", type=int, default=0)
    ap.add_argument("--first", type=int, default=0)
    args = ap.parse_args()
    text = open(args.csv, errors="replace").read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    rows = [r for r in rows if r.get("Metric Name", "").startswith("gpu__time_duration")]
    if args.first:
        rows = rows[: args.first]
    if args.last:
        rows = rows[-args.last:]
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0}.get(unit, 1.0)
        agg[name][0] += 1
        agg[name][1] += ns
        total += ns
    print(f"launches: {len(rows)}   total device time: {total / 1e6:.3f} ms (serialised, cold cache — compare shares)\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.1f} |")


if __name__ == "__main__":
    main()
