#!/usr/bin/env python3
"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv` (one column per metric)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct", "gpu__dram_throughput",
                        "sm__pipe_tensor", "sm__warps_active", "launch__registers", "sm__throughput.avg.pct", "launch__occupancy",
                        "smsp__issue_active.avg.pct", "l1tex__data_bank_conflicts", "smsp__average_warp", "lts__t_sector_hit_rate",
                        "sm__inst_executed_pipe_tensor", "smsp__warp_issue_stalled", "launch__grid_size", "launch__shared_mem"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[4][:80])
    for h, u, v in zip(hdr, units, r):
        if any(p in h for p in pats):
            print(f"  {h} [{u}] = {v}")
