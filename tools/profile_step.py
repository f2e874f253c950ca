#!/usr/bin/env python3
"""Per-kernel GPU time of one eager training step via torch.profiler (CUPTI activity records — no
kernel replay, so it is cheap compared with ncu).  Usage: python tools/profile_step.py [--batch 8]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402
import csts_oracle as O  # noqa: E402
from csts_b200.host.build import build_model  # noqa: E402
from csts_b200.host.train_step import construct_optimizer, train_step  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--rows", type=int, default=45)
    args = ap.parse_args()
    cfg = bench.make_cfg(1)
    model = build_model(cfg)
    model.train()
    opt = construct_optimizer(model, cfg)
    v, a, h = (t.cuda() for t in O.synthetic_batch(args.batch, seed=1))
    for _ in range(3):
        train_step(cfg, model, opt, [v], a, h)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        train_step(cfg, model, opt, [v], a, h)
        torch.cuda.synchronize()
    evs = [e for e in prof.key_averages() if e.device_time_total > 0]
    evs.sort(key=lambda e: -e.device_time_total)
    total = sum(e.device_time_total for e in evs)
    print(f"total GPU kernel time {total / 1e3:.3f} ms over {sum(e.count for e in evs)} launches")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for e in evs[: args.rows]:
        print(f"| `{e.key[:100]}` | {e.count} | {e.device_time_total / 1e3:.3f} | {100 * e.device_time_total / total:.1f}% | "
              f"{e.device_time_total / e.count:.1f} |")


if __name__ == "__main__":
    main()
