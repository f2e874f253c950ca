#!/usr/bin/env python3
"""One markdown row per launch of an `ncu --set full` report (read with `ncu -i X --page raw --csv`).

    python tools/ncu_full_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>_full.md"""
import csv
import subprocess
import sys

COLS = [
    ("us", "gpu__time_duration.sum", 1.0),
    ("grid", "launch__grid_size", 1.0),
    ("regs", "launch__registers_per_thread", 1.0),
    ("smem KB", "launch__shared_mem_per_block_allocated", 1.0),
    ("DRAM rd MB", "dram__bytes_read.sum", 1.0),
    ("DRAM wr MB", "dram__bytes_write.sum", 1.0),
    ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("L2 hit %", "lts__t_sector_hit_rate.pct", 1.0),
    ("warps/cyc", "sm__warps_active.avg.per_cycle_active", 1.0),
    ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
]
TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
TO_KB = {"byte/block": 1e-3, "Kbyte/block": 1.0}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"`ncu --set full --clock-control none` capture: {rep.split('/')[-1]} (per launch; profiler-replayed, cold cache)\n")
    print("| kernel | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        cells = []
        for label, key, _ in COLS:
            i = idx.get(key)
            if i is None or r[i] == "":
                cells.append("")
                continue
            v = float(r[i].replace(",", ""))
            u = units[i]
            if "MB" in label:
                v *= TO_MB.get(u, 1.0)
            elif label == "us":
                v *= TO_US.get(u, 1.0)
            elif "KB" in label:
                v *= TO_KB.get(u, 1.0)
            cells.append(f"{v:.1f}" if v < 1000 and v != int(v) else f"{v:.0f}")
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
