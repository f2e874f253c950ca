#!/bin/bash
# A/B of the factored frame-pool gradient exchange under data parallelism: $1 = ranks
N=${1:-2}
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-second-mode --no-gpu-reference --no-cpu-baseline"
for i in 1 2; do
  timeout 240 $B > gpurun_out/abdp${N}_fact_$i.json 2> gpurun_out/abdp${N}_fact_$i.err
  CSTS_FACTORED_WGRAD=0 timeout 240 $B > gpurun_out/abdp${N}_allreduce_$i.json 2> gpurun_out/abdp${N}_allreduce_$i.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/abdp${N}_*_?.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"], 3), round(d["value"], 1), d.get("dp_check", {}).get("rel"), d.get("dp_check", {}).get("run_to_run_rel"))
    except Exception as e:
        print(f, "failed", e)
PY
