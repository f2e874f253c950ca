#!/usr/bin/env python3
"""Launch a few named GEMM problems of the training step (for `ncu --set full` captures and quick timings).

    python benchmarks/probe_gemm.py [--reps 3] [--time] name[,name...]

names: fc1_dec4 (262144x384x192 GELU), fc1_enc (8192x1536x384 GELU), fc1_dec2 (32768x1536x768 GELU), qkv_s4 (8192x2304x768), dgelu_dec3 (131072x384x192 times-Z, MN-major B),
       fc2_enc (8192x384x1536 + residual f32), qkv_big (131072x576x192), wgrad_fc1 (1536x384 over 8192 tokens, split-K),
       qk_softmax (1024x256x96 x32 heads), pv (1024x96x256 x32)
"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from csts_b200 import kernels as K  # noqa: E402

dev = "cuda"


def rnd(*shape, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(dtype)


def make(name):
    if name in ("fc1_dec4", "fc1_enc", "fc1_dec3", "fc1_dec2"):
        M, N, Kd = {"fc1_dec4": (262144, 384, 192), "fc1_enc": (8192, 1536, 384), "fc1_dec3": (131072, 768, 384), "fc1_dec2": (32768, 1536, 768)}[name]
        A, B, bias = rnd(M, Kd), rnd(N, Kd, scale=0.05), torch.randn(N, device=dev)
        Z, out = torch.empty(M, N, dtype=torch.bfloat16, device=dev), torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        return lambda **kw: K.gemm(A, B, M=M, N=N, K=Kd, bias=bias, act=1, Z=Z, out=out, **kw), 2.0 * M * N * Kd, 2.0 * (M * Kd + N * Kd + 2 * M * N)
    if name == "dgelu_dec3":
        M, N, Kd = 131072, 384, 192
        A, B = rnd(M, Kd), rnd(Kd, N, scale=0.05)
        Z, out = rnd(M, N), torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        return lambda **kw: K.gemm(A, B, M=M, N=N, K=Kd, b_kmajor=False, act=2, Z=Z, out=out, **kw), 2.0 * M * N * Kd, 2.0 * (M * Kd + N * Kd + 2 * M * N)
    if name == "fc2_enc":
        M, N, Kd = 8192, 384, 1536
        A, B, bias = rnd(M, Kd), rnd(N, Kd, scale=0.05), torch.randn(N, device=dev)
        res, out = torch.randn(M, N, device=dev), torch.empty(M, N, device=dev)
        return lambda **kw: K.gemm(A, B, M=M, N=N, K=Kd, bias=bias, residual=res, out=out, **kw), 2.0 * M * N * Kd, 2.0 * (M * Kd + N * Kd) + 8.0 * M * N
    if name in ("qkv_big", "qkv_s4"):
        M, N, Kd = (131072, 576, 192) if name == "qkv_big" else (8192, 2304, 768)
        A, B, bias = rnd(M, Kd), rnd(N, Kd, scale=0.05), torch.randn(N, device=dev)
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        return lambda **kw: K.gemm(A, B, M=M, N=N, K=Kd, bias=bias, out=out, **kw), 2.0 * M * N * Kd, 2.0 * (M * Kd + N * Kd + M * N)
    if name == "wgrad_fc1":
        Mo, No, T = 1536, 384, 8192
        dY, X = rnd(T, Mo), rnd(T, No)
        dw, db = torch.zeros(Mo, No, device=dev), torch.zeros(Mo, device=dev)
        return (lambda **kw: K.gemm(dY, X, M=Mo, N=No, K=T, a_kmajor=False, b_kmajor=False, lda=Mo, ldb=No, out=dw, out_is_zero=True, split_k=-1,
                                    rowsum=db, **kw)), 2.0 * Mo * No * T, 2.0 * T * (Mo + No) + 4.0 * Mo * No
    if name in ("qk_softmax", "pv"):
        B_, h, Lq, Lk, d = 8, 4, 1024, 256, 96
        q, k = rnd(B_, h, Lq, d), rnd(B_, h, Lk, d)
        P = torch.empty(B_, h, Lq, Lk, dtype=torch.bfloat16, device=dev)
        if name == "qk_softmax":
            return (lambda **kw: K.gemm(q, k, M=Lq, N=Lk, K=d, out=P, ldc=Lk, alpha=d ** -0.5, act=3, batch=(B_, h), sA=(h * Lq * d, Lq * d),
                                        sB=(h * Lk * d, Lk * d), sC=(h * Lq * Lk, Lq * Lk), **kw)), 2.0 * B_ * h * Lq * Lk * d, 2.0 * B_ * h * (Lq * d + Lk * d + Lq * Lk)
        o = torch.empty(B_ * Lq, h * d, dtype=torch.bfloat16, device=dev)
        P.normal_()
        return (lambda **kw: K.gemm(P, k, M=Lq, N=d, K=Lk, lda=Lk, b_kmajor=False, ldb=d, out=o, ldc=h * d, batch=(B_, h), sA=(h * Lq * Lk, Lq * Lk),
                                    sB=(h * Lk * d, Lk * d), sC=(Lq * h * d, d), **kw)), 2.0 * B_ * h * Lq * Lk * d, 2.0 * B_ * h * (Lq * d + Lk * d + Lq * Lk)
    raise SystemExit(f"unknown problem {name}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--tile-n", type=int, default=0)
    ap.add_argument("--ctas", type=int, default=0)
    args = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name in args.names.split(","):
        fn, flops, byts = make(name)
        kw = dict(tile_n=args.tile_n, ctas=args.ctas)
        fn(**kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn(**kw)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        if args.time:
            us = min(ts)
            print(f"{name}: {us:.1f} us  {flops / us / 1e6:.0f} TF/s  {byts / us / 1e3:.0f} GB/s (cold L2, incl. launch)")


if __name__ == "__main__":
    main()
