#!/usr/bin/env python3
"""Where the data-parallel overhead of the training step goes (timing diagnostic, launched like bench.py):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
        benchmarks/dp_breakdown.py [--steps 20]

Times the captured step (batch 8 per rank, max over ranks, L2 flushed between replays) in four variants:
  full        the product path: bucketed in-place all-reduce + factored frame-pool gradients
  allreduce   the frame-pool gradients all-reduced like every other tensor (CSTS_FACTORED_WGRAD=0)
  no_reduce   the bucket all-reduces skipped (factor all-gathers and the NCE all-gather remain)
  compute     no gradient exchange at all (only the NCE all-gather of the loss remains)
The last two produce wrong gradients on purpose: they isolate what the exchange costs beyond the arithmetic.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    import bench
    import csts_oracle as O
    from csts_b200.host import distributed as du
    from csts_b200.host.build import build_model
    from csts_b200.host.train_step import GraphedTrainStep, construct_optimizer, make_grad_scaler

    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = bench.make_cfg(world, "fp16")
    video, audio, hm = (t.to(dev) for t in O.synthetic_batch(bench.BATCH_PER_GPU, seed=100 + rank))
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
    real_hook = du.OverlappedGradSync._hook

    def counting_hook(self, param):              # bucket bookkeeping without the collective
        if self.enabled:
            self.arena.adopt(param)
            self.pending[self.group_of[id(param)]] -= 1

    out, keep = {}, []
    for name, factored, hook in (("full", "1", real_hook), ("allreduce", "0", real_hook), ("no_reduce", "1", counting_hook),
                                 ("compute", "0", counting_hook)):
        os.environ["CSTS_FACTORED_WGRAD"] = factored
        du.OverlappedGradSync._hook = hook
        torch.manual_seed(cfg.RNG_SEED)
        model = build_model(cfg, ddp=False)
        model.train()
        opt = construct_optimizer(model, cfg, capturable=True, fused_clip=True)
        g = GraphedTrainStep(cfg, model, opt, video, audio, hm, scaler=make_grad_scaler(cfg))
        for _ in range(5):
            g(None, None, None)
        dist.barrier()
        torch.cuda.synchronize()
        evs = []
        for _ in range(args.steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            g(None, None, None)
            e.record()
            evs.append((s, e))
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([sum(s.elapsed_time(e) for s, e in evs) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = round(t.item(), 3)
        keep.append((g, opt, model))             # captured NCCL kernels stay referenced until the process exits
    du.OverlappedGradSync._hook = real_hook
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(json.dumps({"n_gpus": world, "ms_per_step": out, "unit": "ms", "steps": args.steps}), flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
