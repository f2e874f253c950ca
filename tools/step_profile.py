#!/usr/bin/env python3
"""Per-entry-point CUDA-event timing of one eager training step (batch 8): every C-ABI call is bracketed by an event pair
while the GPU is parked behind a spin kernel, so the events bracket kernels that run back to back with warm caches (what
the graph replay sees), unlike ncu's serialised cold-cache launch list.   python tools/step_profile.py [--precision bf16]"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--detail", action="store_true", help="group by entry point AND integer arguments (shapes)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "step_profile.json"))
    args = ap.parse_args()
    import torch
    import csts_oracle as O
    from csts_b200 import _lib
    from csts_b200.host.build import build_model
    from csts_b200.host.config import get_cfg
    from csts_b200.host.train_step import construct_optimizer, make_grad_scaler, train_step
    dev = torch.device("cuda", 0)
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(ROOT, "configs", "Ego4D", "CSTS_Ego4D_Gaze_Forecast.yaml"))
    cfg.merge_from_list(["NUM_GPUS", 1, "MODEL.LOSS_FUNC", "kldiv+egonce", "TRAIN.MIXED_PRECISION", args.precision == "fp16"])
    torch.manual_seed(0)
    model = build_model(cfg)
    model.train()
    model._wc.fork_backward = False
    model._wc.parallel_audio = False
    opt = construct_optimizer(model, cfg, capturable=True, fused_clip=True)
    scaler = make_grad_scaler(cfg)
    video, audio, hm = (t.to(dev) for t in O.synthetic_batch(args.batch, seed=1))
    for _ in range(3):
        train_step(cfg, model, opt, [video], audio, hm, scaler=scaler)
    torch.cuda.synchronize()
    _lib.PROFILE = []
    torch.cuda._sleep(int(0.12 * 1.9e9))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    train_step(cfg, model, opt, [video], audio, hm, scaler=scaler)
    e.record()
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    fam = collections.OrderedDict()
    for name, e0, e1 in prof:
        if not args.detail:
            name = name.split("(")[0]
        f = fam.setdefault(name, [0, 0.0])
        f[0] += 1
        f[1] += e0.elapsed_time(e1)
    tot = sum(v[1] for v in fam.values())
    rows = sorted(fam.items(), key=lambda kv: -kv[1][1])
    print(f"entry points: {len(prof)} calls, {tot:.2f} ms inside event pairs (step wall incl. torch ops: {s.elapsed_time(e):.2f} ms)")
    for n, (c, ms) in rows:
        print(f"{n:60s} {c:5d} calls {ms:8.3f} ms {100 * ms / tot:5.1f}%  avg {1e3 * ms / c:7.1f} us")
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"total_ms": tot, "rows": [(n, c, ms) for n, (c, ms) in rows]}, f, indent=1)


if __name__ == "__main__":
    main()
