"""BASELINE.json configs[4]: MultiScaleAttention / fusion attention microbenchmark at every attention
site of CSTS, B = 8: QK^T (tcgen05, batched strided) -> softmax -> PV forward, and
the four backward products + softmax backward.  Reports time, TFLOP/s of the 4*B*h*Lq*Lk*d algorithmic
FLOPs (x2.5 for backward) and the fraction of the measured bf16 tensor peak."""
import json
import sys

import torch

sys.path.insert(0, ".")
from csts_b200 import kernels as K  # noqa: E402

B = 8
SITES = [  # name, heads, d, Lq, Lk, q_strided_in_qkv
    ("blocks.0", 1, 96, 16384, 256, True), ("blocks.1", 2, 96, 4096, 1024, False), ("blocks.2", 2, 96, 4096, 256, True),
    ("blocks.3", 4, 96, 1024, 1024, False), ("blocks.4-13", 4, 96, 1024, 256, True), ("blocks.14", 8, 96, 256, 1024, False),
    ("blocks.15", 8, 96, 256, 256, True), ("spatial_fusion", 8, 96, 260, 260, True), ("decode_block1", 8, 96, 1024, 64, False),
    ("decode_block2", 4, 192, 4096, 64, False), ("decode_block3", 4, 96, 16384, 64, False), ("decode_block4", 2, 96, 32768, 64, False),
]


def graph_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def main():
    dev, peak = "cuda", 1402.0
    for name, h, d, Lq, Lk, strided in SITES:
        Cn = h * d
        ldS = (Lk + 7) // 8 * 8
        if strided:
            qbuf = torch.randn(B, Lq, 3, h, d, device=dev).to(torch.bfloat16)
            q_ld, q_s = 3 * Cn, (Lq * 3 * Cn, d)
        else:
            qbuf = torch.randn(B, h, Lq, d, device=dev).to(torch.bfloat16)
            q_ld, q_s = d, (h * Lq * d, Lq * d)
        k = torch.randn(B, h, Lk, d, device=dev).to(torch.bfloat16)
        v = torch.randn(B, h, Lk, d, device=dev).to(torch.bfloat16)
        S = torch.empty(B, h, Lq, ldS, dtype=torch.float32, device=dev)
        o = torch.empty(B, Lq, Cn, dtype=torch.bfloat16, device=dev)
        do = torch.randn(B, Lq, Cn, device=dev).to(torch.bfloat16)
        dq, dk, dv = torch.empty_like(qbuf), torch.empty_like(k), torch.empty_like(v)
        dP = torch.empty_like(S)
        sP, kv_s = (h * Lq * ldS, Lq * ldS), (h * Lk * d, Lk * d)
        mask = dict(mask_hw=64, mask_t=4) if name == "spatial_fusion" else {}
        state = {}

        def fwd():
            K.gemm(qbuf, k, M=Lq, N=Lk, K=d, lda=q_ld, ldb=d, out=S, ldc=ldS, alpha=d ** -0.5, batch=(B, h), sA=q_s, sB=kv_s, sC=sP)
            state["P"] = K.softmax_fwd(S, Lk, ldS, nq=Lq, **mask)
            K.gemm(state["P"], v, M=Lq, N=d, K=Lk, lda=ldS, b_kmajor=False, ldb=d, out=o, ldc=Cn, batch=(B, h), sA=sP, sB=kv_s, sC=(Lq * Cn, d))

        def bwd():
            P = state["P"]
            K.gemm(P, do, M=Lk, N=d, K=Lq, a_kmajor=False, lda=ldS, b_kmajor=False, ldb=Cn, out=dv, ldc=d, batch=(B, h), sA=sP,
                   sB=(Lq * Cn, d), sC=kv_s)
            K.gemm(do, v, M=Lq, N=Lk, K=d, lda=Cn, ldb=d, out=dP, ldc=ldS, batch=(B, h), sA=(Lq * Cn, d), sB=kv_s, sC=sP)
            dS = K.softmax_bwd(P, dP, Lk, d ** -0.5)
            K.gemm(dS, k, M=Lq, N=d, K=Lk, lda=ldS, b_kmajor=False, ldb=d, out=dq, ldc=q_ld, batch=(B, h), sA=sP, sB=kv_s, sC=q_s)
            K.gemm(dS, qbuf, M=Lk, N=d, K=Lq, a_kmajor=False, lda=ldS, b_kmajor=False, ldb=q_ld, out=dk, ldc=d, batch=(B, h), sA=sP,
                   sB=q_s, sC=kv_s)

        t_f = graph_time(fwd)
        t_b = graph_time(bwd)
        fl = 4.0 * B * h * Lq * Lk * d
        print(json.dumps({"site": name, "heads": h, "d": d, "Lq": Lq, "Lk": Lk, "fwd_ms": round(t_f, 4), "bwd_ms": round(t_b, 4),
                          "fwd_tflops": round(fl / t_f / 1e9, 1), "bwd_tflops": round(2.0 * fl / t_b / 1e9, 1),
                          "fwd_frac_of_bf16_peak": round(fl / t_f / 1e9 / peak, 3)}), flush=True)


if __name__ == "__main__":
    main()
