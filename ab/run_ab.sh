#!/bin/bash
# A/B of GEMM work-item scheduling on one box: previous library / dynamic / static
B="python bench.py --steps 20 --warmup 5 --no-second-mode --no-gpu-reference --no-cpu-baseline"
for i in 1 2; do
  CSTS_B200_LIB=/root/repo/ab/libcsts_old.so timeout 200 $B > gpurun_out/ab_old_$i.json 2> gpurun_out/ab_old_$i.err
  timeout 200 $B > gpurun_out/ab_dyn_$i.json 2> gpurun_out/ab_dyn_$i.err
  CSTS_GEMM_STATIC=1 timeout 200 $B > gpurun_out/ab_static_$i.json 2> gpurun_out/ab_static_$i.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"], 3), round(d["roofline"]["kernel_ms_per_step"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
