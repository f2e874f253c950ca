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

        # the path host/block.py takes at this site: fused whole-row softmax epilogue (<= 256 keys), two-pass logsumexp +
        # exp epilogue (> 256 keys), or f32 scores + masked softmax kernel (spatial fusion)
        fused = Lk <= 256 and Lq >= 64 and name != "spatial_fusion"
        two_pass = not fused and name != "spatial_fusion"
        scale = d ** -0.5
        P = torch.empty(B, h, Lq, ldS, dtype=torch.bfloat16, device=dev)
        dS = torch.empty_like(P)
        lse = torch.empty(B, h, Lq, dtype=torch.float32, device=dev)
        qk = dict(M=Lq, N=Lk, K=d, lda=q_ld, ldb=d, alpha=scale, batch=(B, h), sA=q_s, sB=kv_s)

        def fwd():
            if fused:
                K.gemm(qbuf, k, out=P, ldc=ldS, act=3, sC=sP, **qk)
            elif two_pass:
                K.gemm(qbuf, k, out=lse, ldc=Lk, act=5, sC=(h * Lq, Lq), **qk)
                K.gemm(qbuf, k, out=P, ldc=ldS, act=6, rowvec=lse, sC=sP, **qk)
            else:
                K.gemm(qbuf, k, out=S, ldc=ldS, sC=sP, **qk)
                state["P"] = K.softmax_fwd(S, Lk, ldS, nq=Lq, **mask)
            K.gemm(state.get("P", P), v, M=Lq, N=d, K=Lk, lda=ldS, b_kmajor=False, ldb=d, out=o, ldc=Cn, batch=(B, h), sA=sP, sB=kv_s, sC=(Lq * Cn, d))

        def bwd():
            Pm = state.get("P", P)
            K.gemm(Pm, do, M=Lk, N=d, K=Lq, a_kmajor=False, lda=ldS, b_kmajor=False, ldb=Cn, out=dv, ldc=d, batch=(B, h), sA=sP,
                   sB=(Lq * Cn, d), sC=kv_s)
            if fused:
                K.gemm(do, v, M=Lq, N=Lk, K=d, lda=Cn, ldb=d, out=dS, ldc=ldS, alpha=scale, act=4, Z=Pm, batch=(B, h), sA=(Lq * Cn, d), sB=kv_s, sC=sP)
                ds = dS
            elif two_pass:
                D = K.rowdot(do.view(B * Lq, Cn), o.view(B * Lq, Cn), B, Lq, h, d)
                K.gemm(do, v, M=Lq, N=Lk, K=d, lda=Cn, ldb=d, out=dS, ldc=ldS, alpha=scale, act=7, Z=Pm, rowvec=D, batch=(B, h), sA=(Lq * Cn, d),
                       sB=kv_s, sC=sP)
                ds = dS
            else:
                K.gemm(do, v, M=Lq, N=Lk, K=d, lda=Cn, ldb=d, out=dP, ldc=ldS, batch=(B, h), sA=(Lq * Cn, d), sB=kv_s, sC=sP)
                ds = K.softmax_bwd(Pm, dP, Lk, scale)
            K.gemm(ds, k, M=Lq, N=d, K=Lk, lda=ldS, b_kmajor=False, ldb=d, out=dq, ldc=q_ld, batch=(B, h), sA=sP, sB=kv_s, sC=q_s)
            K.gemm(ds, qbuf, M=Lk, N=d, K=Lq, a_kmajor=False, lda=ldS, b_kmajor=False, ldb=q_ld, out=dk, ldc=d, batch=(B, h), sA=sP,
                   sB=q_s, sC=kv_s)

        t_f = graph_time(fwd)
        t_b = graph_time(bwd)
        fl = 4.0 * B * h * Lq * Lk * d
        print(json.dumps({"site": name, "heads": h, "d": d, "Lq": Lq, "Lk": Lk, "fwd_ms": round(t_f, 4), "bwd_ms": round(t_b, 4),
                          "fwd_tflops": round(fl / t_f / 1e9, 1), "bwd_tflops": round(2.0 * fl / t_b / 1e9, 1),
                          "fwd_frac_of_bf16_peak": round(fl / t_f / 1e9 / peak, 3), "bwd_frac_of_bf16_peak": round(2.0 * fl / t_b / 1e9 / peak, 3),
                          "path": "fused-softmax epilogue" if fused else ("two-pass (lse + exp epilogue)" if two_pass else "f32 scores + masked softmax")}),
              flush=True)


if __name__ == "__main__":
    main()
