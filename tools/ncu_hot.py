#!/usr/bin/env python3
"""Top stall-sample instructions from `ncu --page source --csv` (SASS view)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out[1:]))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[1:]:
    try:
        data.append((int(r[idx["# Samples"]] or 0), r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for n, r in sorted(data, key=lambda d: -d[0])[:top]:
    st = sorted(((int(r[idx[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{n:6d} {100*n/tot:5.1f}%  {r[idx['Source']][:90]:90s} {st}")
