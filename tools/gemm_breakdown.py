#!/usr/bin/env python3
"""Per-shape breakdown of every GEMM launch of one eager training step (CUDA events around each launch):
time, achieved TFLOP/s and GB/s against the measured peaks.  python tools/gemm_breakdown.py > profiles/..."""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402
import csts_oracle as O  # noqa: E402
from csts_b200 import kernels as K  # noqa: E402
from csts_b200.host.build import build_model  # noqa: E402
from csts_b200.host.train_step import construct_optimizer, train_step  # noqa: E402


def main():
    cfg = bench.make_cfg(1)
    model = build_model(cfg)
    model.train()
    opt = construct_optimizer(model, cfg)
    v, a, h = (t.cuda() for t in O.synthetic_batch(8, seed=1))
    for _ in range(3):
        train_step(cfg, model, opt, [v], a, h)
    torch.cuda.synchronize()
    K.GEMM_PROFILE = []
    for _ in range(3):
        torch.cuda._sleep(int(0.08 * 1.9e9))     # park the GPU so the eager step is fully enqueued before it runs
        train_step(cfg, model, opt, [v], a, h)
    torch.cuda.synchronize()
    prof, K.GEMM_PROFILE = K.GEMM_PROFILE, None
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for ev0, ev1, fl, by, tc, key in prof:
        r = agg[(tc,) + key]
        r[0] += 1
        r[1] += ev0.elapsed_time(ev1)
        r[2] += fl
        r[3] += by
    hbm, tf, _ = bench.peaks()
    total = sum(r[1] for r in agg.values()) / 3
    print(f"GEMM time per step (event-timed, eager): {total:.3f} ms over {sum(r[0] for r in agg.values()) // 3} launches\n")
    print("| kernel | M | N | K | batch | A,B K-major | act | res | out | splitK | launches/step | ms/step | TFLOP/s | GB/s | frac tensor | frac HBM |")
    print("|---|---|---|---|---|---|---|---|---|---|---:|---:|---:|---:|---:|---:|")
    for key, (n, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        tc, M, N, Kd, nb, ak, bk, act, res, cdt, sk = key
        tfl, gbs = fl / ms / 1e9, by / ms / 1e6
        print(f"| {'tcgen05' if tc else 'mma.sync'} | {M} | {N} | {Kd} | {nb} | {ak},{bk} | {act} | {res} | {'bf16' if cdt else 'f32'} | {sk} | "
              f"{n // 3} | {ms / 3:.3f} | {tfl:.0f} | {gbs:.0f} | {tfl / tf:.2f} | {gbs / hbm:.2f} |")


if __name__ == "__main__":
    main()
