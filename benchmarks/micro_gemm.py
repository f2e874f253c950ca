"""GEMM microbenchmark at the CSTS Linear shapes (SURVEY.md App. A.3, B=8): tcgen05 kernel vs the
mma.sync kernel vs torch.matmul (cuBLAS), CUDA-event timed, L2 flushed between iterations."""
import json
import sys

import torch

sys.path.insert(0, ".")
from csts_b200 import kernels as K  # noqa: E402

SHAPES = [  # (M tokens at B=8, K, N)
    (131072, 96, 288), (131072, 96, 384), (131072, 384, 192), (131072, 192, 576), (32768, 192, 768), (32768, 768, 384),
    (32768, 384, 1152), (8192, 384, 1536), (8192, 1536, 384), (8192, 768, 2304), (2048, 768, 3072), (2048, 3072, 768),
    (8192, 768, 3072), (8192, 3072, 768), (32768, 768, 1536), (131072, 384, 768), (131072, 768, 192), (262144, 192, 384),
    (262144, 384, 96),
]


def timeit(fn, flush, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = "cuda"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []
    for M, Kd, N in SHAPES:
        A = torch.randn(M, Kd, device=dev).to(torch.bfloat16)
        B = (torch.randn(N, Kd, device=dev) * 0.05).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        t_tc = timeit(lambda: K.gemm(A, B, M=M, N=N, K=Kd, bias=bias, out=out, backend=2), flush)
        t_mma = timeit(lambda: K.gemm(A, B, M=M, N=N, K=Kd, bias=bias, out=out, backend=1), flush)
        t_lib = timeit(lambda: torch.addmm(bias.to(torch.bfloat16), A, B.t(), out=out), flush)
        flops = 2.0 * M * N * Kd
        byts = 2.0 * (M * Kd + N * Kd + M * N)
        rows.append(dict(M=M, K=Kd, N=N, tc_ms=t_tc, mma_ms=t_mma, cublas_ms=t_lib, tc_tflops=flops / t_tc / 1e9,
                         tc_gbs=byts / t_tc / 1e6, cublas_tflops=flops / t_lib / 1e9))
        print(json.dumps(rows[-1]))
    return rows


if __name__ == "__main__":
    main()
