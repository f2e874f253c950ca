#!/usr/bin/env python3
"""Measure, per GEMM problem of the CSTS training step, which (tile width, CTAs per SM, split-K) the tcgen05 launcher
should use, and write the table the library compiles in (csts_b200/csrc/gemm_tune.inc).

One eager training step at the benchmark configuration (batch 8) is recorded: every `kernels.gemm` call with its real
operands (strides, epilogue inputs, batch layout).  Each distinct problem is then replayed on those operands with every
candidate plan inside a CUDA graph (back-to-back launches, no host gaps) and timed with CUDA events.

    python benchmarks/tune_gemm.py [--batch 8] [--precision bf16] [--out gpurun_out/gemm_tune.json] [--emit]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def key_of(kw):
    b1, b2 = kw.get("batch", (1, 1))
    out = kw.get("out")
    odt = kw.get("out_dtype")
    import torch
    if out is not None:
        c = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[out.dtype]
    elif odt is not None:
        c = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[odt]
    else:
        c = 1
    return (kw["M"], kw["N"], kw["K"], b1 * b2, int(kw.get("a_kmajor", True)), int(kw.get("b_kmajor", True)), kw.get("act", 0),
            min(c, 1), int(kw.get("rowsum") is not None), int(kw.get("split_k", 1) < 0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gemm_tune.json"))
    ap.add_argument("--emit", action="store_true", help="also write csts_b200/csrc/gemm_tune.inc")
    ap.add_argument("--reps", type=int, default=24)
    args = ap.parse_args()
    import torch
    import csts_oracle as O
    from csts_b200 import _lib, kernels as K
    from csts_b200.host.build import build_model
    from csts_b200.host.config import get_cfg
    from csts_b200.host.train_step import compute_loss, make_grad_scaler
    os.environ["CSTS_GEMM_NO_TABLE"] = "1"          # candidates are measured against the model's choice, not an older table
    dev = torch.device("cuda", 0)
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(ROOT, "configs", "Ego4D", "CSTS_Ego4D_Gaze_Forecast.yaml"))
    cfg.merge_from_list(["NUM_GPUS", 1, "MODEL.LOSS_FUNC", "kldiv+egonce", "TRAIN.MIXED_PRECISION", args.precision == "fp16"])
    torch.manual_seed(0)
    model = build_model(cfg)
    model.train()
    model._wc.fork_backward = False
    video, audio, hm = (t.to(dev) for t in O.synthetic_batch(args.batch, seed=1))
    calls = {}
    orig = K.gemm

    def recorder(A, B, **kw):
        r = orig(A, B, **kw)
        k = key_of(kw)
        if k not in calls and _lib.load().csts_gemm_backend is not None:
            kw2 = dict(kw)
            if kw2.get("out") is None:
                kw2["out"] = r
            calls[k] = (A, B, kw2, [1])
        elif k in calls:
            calls[k][3][0] += 1
        return r

    scaler = make_grad_scaler(cfg)
    K.gemm = recorder
    loss, _, _, _ = compute_loss(cfg, model, [video], audio, hm)
    (scaler.scale(loss) if scaler.is_enabled() else loss).backward()
    K.gemm = orig
    torch.cuda.synchronize()
    print(f"{len(calls)} distinct GEMM problems, {sum(c[3][0] for c in calls.values())} launches per step", file=sys.stderr)

    def time_plan(A, B, kw, tile_n, ctas, split):
        kw = dict(kw)
        kw.update(tile_n=tile_n, ctas=ctas, backend=2)
        if split is not None:
            kw["split_k"] = split
        try:
            orig(A, B, **kw)
            torch.cuda.synchronize()
        except RuntimeError:
            return None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(args.reps):
                orig(A, B, **kw)
        g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        return 1e3 * s.elapsed_time(e) / args.reps

    lib = _lib.load()
    rows = []
    for k, (A, B, kw, cnt) in sorted(calls.items(), key=lambda it: -it[1][3][0]):
        M, N, Kd, nb, ak, bk, act, c16, rs, auto = k
        probe = dict(kw)
        probe["backend"] = 0
        # only problems the tcgen05 kernel takes
        ga = K.build_gemm_args(A, B, **probe)
        if lib.csts_gemm_backend(C.byref(ga)) != 2:
            continue
        bn0, ct0, sp0 = C.c_int(), C.c_int(), C.c_int()
        lib.csts_gemm_plan(C.byref(ga), C.byref(bn0), C.byref(ct0), C.byref(sp0))
        base = time_plan(A, B, kw, bn0.value, ct0.value, sp0.value if auto else None)
        if act in (3, 4):
            tiles = [bn0.value]
        elif rs:
            tiles = [96] if N <= 96 else [192]
        else:
            pad = min(-(-N // t) * t for t in (96, 128, 192, 256))
            tiles = [t for t in (96, 128, 192, 256) if -(-N // t) * t == pad]
        kblocks = -(-Kd // 64)
        splits = [None]
        if auto:
            splits = [s for s in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64) if s == 1 or kblocks // s >= 2]
        best = (base, bn0.value, ct0.value, sp0.value)
        allr = []
        for t in tiles:
            for ct in (1, 2):
                if ct == 2 and (t == 256 or act in (3, 4)):
                    continue
                for s in splits:
                    us = time_plan(A, B, kw, t, ct, s)
                    if us is None:
                        continue
                    allr.append((round(us, 2), t, ct, s))
                    if us < best[0] * 0.97:
                        best = (us, t, ct, s if s is not None else 1)
        rows.append(dict(key=list(k), launches=cnt[0], model_plan=[bn0.value, ct0.value, sp0.value], model_us=round(base, 2),
                         best_plan=[best[1], best[2], best[3]], best_us=round(best[0], 2), all=sorted(allr)[:6]))
        print(json.dumps(rows[-1]), flush=True)
    tot0 = sum(r["model_us"] * r["launches"] for r in rows)
    tot1 = sum(r["best_us"] * r["launches"] for r in rows)
    print(f"step total: model plan {tot0 / 1e3:.2f} ms, tuned {tot1 / 1e3:.2f} ms over {sum(r['launches'] for r in rows)} launches", file=sys.stderr)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(dict(batch=args.batch, precision=args.precision, model_ms=tot0 / 1e3, tuned_ms=tot1 / 1e3, rows=rows), f, indent=0)
    if args.emit:
        emit(rows, os.path.join(ROOT, "csts_b200", "csrc", "gemm_tune.inc"), args)


def emit(rows, path, args):
    with open(path, "w") as f:
        f.write("// Measured launch plans of the tcgen05 GEMM for the problems of one CSTS training step (batch %d per GPU),\n"
                "// written by benchmarks/tune_gemm.py on a B200: {M, N, K, batch, a_kmajor, b_kmajor, act, c_16bit, rowsum, auto_split,\n"
                "// tile_n, ctas, splits}.  Problems that are not listed use the launcher's cost model.\n" % args.batch)
        for r in rows:
            if r["best_plan"] == r["model_plan"]:
                continue
            f.write("{%s, %d, %d, %d},\n" % (", ".join(str(v) for v in r["key"]), *r["best_plan"]))


if __name__ == "__main__":
    main()
