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


def timeit(make_launch, sets, reps=4):
    """GPU time per launch with no host overhead and a cold L2: `sets` operand sets (together larger than
    L2) are cycled, the whole sequence is captured into a CUDA graph and replayed between two events."""
    launches = [make_launch(i) for i in range(sets)]
    for fn in launches:
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for fn in launches:
                fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (reps * sets)


def main():
    dev = "cuda"
    rows = []
    for M, Kd, N in SHAPES:
        byts = 2.0 * (M * Kd + N * Kd + M * N)
        sets = max(2, int(300e6 // byts) + 1)
        As = [torch.randn(M, Kd, device=dev).to(torch.bfloat16) for _ in range(sets)]
        Bs = [(torch.randn(N, Kd, device=dev) * 0.05).to(torch.bfloat16) for _ in range(sets)]
        outs = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(sets)]
        bias = torch.randn(N, device=dev)
        bias16 = bias.to(torch.bfloat16)
        t_tc = timeit(lambda i: (lambda: K.gemm(As[i], Bs[i], M=M, N=N, K=Kd, bias=bias, out=outs[i], backend=2)), sets)
        t_mma = timeit(lambda i: (lambda: K.gemm(As[i], Bs[i], M=M, N=N, K=Kd, bias=bias, out=outs[i], backend=1)), sets)
        t_lib = timeit(lambda i: (lambda: torch.addmm(bias16, As[i], Bs[i].t(), out=outs[i])), sets)
        # weight-gradient layout: dW[N, Kd] = dY[M, N]^T . X[M, Kd] (contraction over the M tokens)
        dW = torch.empty(N, Kd, dtype=torch.float32, device=dev)
        split = max(1, min(296 // (((N + 127) // 128) * ((Kd + 127) // 128)), M // 512))
        t_wg = timeit(lambda i: (lambda: K.gemm(outs[i], As[i], M=N, N=Kd, K=M, a_kmajor=False, b_kmajor=False, out=dW, split_k=split,
                                                 backend=2)), sets)
        flops = 2.0 * M * N * Kd
        rows.append(dict(M=M, K=Kd, N=N, tc_us=1e3 * t_tc, mma_us=1e3 * t_mma, cublas_us=1e3 * t_lib, wgrad_tc_us=1e3 * t_wg,
                         tc_tflops=flops / t_tc / 1e9, tc_gbs=byts / t_tc / 1e6, cublas_tflops=flops / t_lib / 1e9,
                         wgrad_tflops=flops / t_wg / 1e9, hbm_floor_us=byts / 6.5558e6, tensor_floor_us=flops / 1402e6))
        print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in rows[-1].items()}), flush=True)
    return rows


if __name__ == "__main__":
    main()
