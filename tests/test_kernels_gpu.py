"""Kernel-level parity (GPU): every exported kernel of libcsts_b200.so against the plain fp32 torch
expression of the reference lines it replaces.  Tolerances are stated per test; bf16 storage
implies ~2^-8 relative rounding on inputs/outputs."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# the torch expressions are the fp32 reference: no TF32 in cuDNN convolutions / cuBLAS matmuls
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

dev = "cuda"


def K():
    from csts_b200 import kernels
    return kernels


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def bf(x):
    return x.to(torch.bfloat16)


def hf(x):
    return x.to(torch.float16)


# ------------------------------------------------------------------------------------------------ GEMM
GEMM_SHAPES = [
    # M, N, K
    (256, 96, 96), (1024, 288, 96), (384, 192, 384), (2080, 768, 768), (128, 2304, 768),
    (300, 384, 1536), (4096, 96, 192), (512, 3072, 768), (1000, 576, 192), (128, 96, 448),
]


@pytest.mark.parametrize("backend", [1, 2])
@pytest.mark.parametrize("M,N,Kd", GEMM_SHAPES)
def test_gemm_linear_plain(backend, M, N, Kd):
    k = K()
    g = torch.Generator(device="cpu").manual_seed(M + N + Kd)
    A = bf(torch.randn(M, Kd, generator=g)).to(dev)
    B = bf(torch.randn(N, Kd, generator=g) * 0.05).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    ref = A.float() @ B.float().t() + bias
    out = k.gemm(A, B, M=M, N=N, K=Kd, bias=bias, out_dtype=torch.float32, backend=backend)
    assert rel_err(out, ref) < 2e-5, rel_err(out, ref)
    out16 = k.gemm(A, B, M=M, N=N, K=Kd, bias=bias, out_dtype=torch.bfloat16, backend=backend)
    assert rel_err(out16, ref) < 4e-3


@pytest.mark.parametrize("backend", [1, 2])
def test_gemm_epilogues(backend):
    k = K()
    M, N, Kd = 640, 384, 192
    g = torch.Generator(device="cpu").manual_seed(5)
    A = bf(torch.randn(M, Kd, generator=g)).to(dev)
    B = bf(torch.randn(N, Kd, generator=g) * 0.1).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev)
    pre = A.float() @ B.float().t() + bias
    # GELU; Z receives GELU'(pre-activation) for the backward epilogue
    Z = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    h = k.gemm(A, B, M=M, N=N, K=Kd, bias=bias, act=1, Z=Z, backend=backend)
    pr = pre.clone().requires_grad_(True)
    (gpre,) = torch.autograd.grad(F.gelu(pr).sum(), pr)
    assert rel_err(Z, gpre) < 4e-3
    assert rel_err(h, F.gelu(pre)) < 5e-3
    # residual, f32 out
    o = k.gemm(A, B, M=M, N=N, K=Kd, bias=bias, residual=res, out_dtype=torch.float32, backend=backend)
    assert rel_err(o, pre + res) < 2e-5
    # broadcast residual rows (pos-embed style)
    o = k.gemm(A, B, M=M, N=N, K=Kd, bias=bias, residual=res[:128].contiguous(), res_mod=128, out_dtype=torch.float32, backend=backend)
    assert rel_err(o, pre + res[:128].repeat(5, 1)) < 2e-5
    # accumulate
    acc = res.clone()
    k.gemm(A, B, M=M, N=N, K=Kd, out=acc, accumulate=True, backend=backend)
    assert rel_err(acc, res + A.float() @ B.float().t()) < 2e-5
    # times Z (GELU backward; the tcgen05 kernel pairs activations with bf16 outputs only)
    o = k.gemm(A, B, M=M, N=N, K=Kd, act=2, Z=Z, out_dtype=torch.float32 if backend == 1 else torch.bfloat16, backend=backend)
    assert rel_err(o, (A.float() @ B.float().t()) * Z.float()) < (1e-4 if backend == 1 else 4e-3)


@pytest.mark.parametrize("backend", [1, 2])
@pytest.mark.parametrize("ak,bk", [(True, True), (True, False), (False, False), (False, True)])
def test_gemm_generic_layouts_batched(ak, bk, backend):
    k = K()
    if backend == 2 and (not ak) and bk:
        pytest.skip("(MN-major A, K-major B) does not occur on the path; generic kernel only")
    b1, b2, M, N, Kd = 2, 3, 260, 96, 264
    Mp = 264                                   # M-major storage needs a leading dim that is a multiple of 8
    g = torch.Generator(device="cpu").manual_seed(11)
    A = bf(torch.randn(b1, b2, M, Kd, generator=g)).to(dev)
    B = bf(torch.randn(b1, b2, Kd, N, generator=g)).to(dev)
    ref = A.float() @ B.float()
    if ak:
        As, lda, sA = A, Kd, (b2 * M * Kd, M * Kd)
    else:
        As = torch.zeros(b1, b2, Kd, Mp, dtype=torch.bfloat16, device=dev)
        As[..., :M] = A.transpose(-1, -2)
        lda, sA = Mp, (b2 * Kd * Mp, Kd * Mp)
    Bs = B.transpose(-1, -2).contiguous() if bk else B
    out = torch.empty(b1, b2, M, N, dtype=torch.float32, device=dev)
    k.gemm(As, Bs, M=M, N=N, K=Kd, a_kmajor=ak, b_kmajor=bk, lda=lda, out=out, batch=(b1, b2),
           sA=sA, sB=(b2 * N * Kd, N * Kd), sC=(b2 * M * N, M * N), alpha=0.5, backend=backend)
    assert rel_err(out, 0.5 * ref) < 2e-5
    out16 = torch.empty(b1, b2, M, N, dtype=torch.bfloat16, device=dev)
    k.gemm(As, Bs, M=M, N=N, K=Kd, a_kmajor=ak, b_kmajor=bk, lda=lda, out=out16, batch=(b1, b2),
           sA=sA, sB=(b2 * N * Kd, N * Kd), sC=(b2 * M * N, M * N), backend=backend)
    assert rel_err(out16, ref) < 4e-3


@pytest.mark.parametrize("backend", [1, 2])
@pytest.mark.parametrize("Lq,Lk,heads", [(256, 64, 2), (512, 256, 2), (260, 260, 8), (128, 1024, 2)])
def test_gemm_attention_shapes(backend, Lq, Lk, heads):
    """The six attention products on strided (B, N, 3, heads, d) / (B, heads, L, ld) operands."""
    k = K()
    B, d = 2, 96
    Cn = heads * d
    g = torch.Generator(device="cpu").manual_seed(Lq + Lk)
    qkv = bf(torch.randn(B, Lq, 3, heads, d, generator=g)).to(dev)                   # q strided inside qkv
    kk = bf(torch.randn(B, heads, Lk, d, generator=g)).to(dev)
    vv = bf(torch.randn(B, heads, Lk, d, generator=g)).to(dev)
    ldS = (Lk + 7) // 8 * 8
    q = qkv[:, :, 0].permute(0, 2, 1, 3).float()
    S = torch.empty(B, heads, Lq, ldS, dtype=torch.float32, device=dev)
    k.gemm(qkv, kk, M=Lq, N=Lk, K=d, lda=3 * Cn, ldb=d, out=S, ldc=ldS, alpha=0.25, batch=(B, heads),
           sA=(Lq * 3 * Cn, d), sB=(heads * Lk * d, Lk * d), sC=(heads * Lq * ldS, Lq * ldS), backend=backend)
    Sref = 0.25 * q @ kk.float().transpose(-1, -2)
    assert rel_err(S[..., :Lk], Sref) < 2e-5
    P = torch.zeros(B, heads, Lq, ldS, dtype=torch.bfloat16, device=dev)
    P[..., :Lk] = bf(Sref.softmax(-1))
    o = torch.empty(B, Lq, Cn, dtype=torch.bfloat16, device=dev)
    k.gemm(P, vv, M=Lq, N=d, K=Lk, lda=ldS, b_kmajor=False, ldb=d, out=o, ldc=Cn, batch=(B, heads),
           sA=(heads * Lq * ldS, Lq * ldS), sB=(heads * Lk * d, Lk * d), sC=(Lq * Cn, d), backend=backend)
    oref = (P[..., :Lk].float() @ vv.float()).permute(0, 2, 1, 3).reshape(B, Lq, Cn)
    assert rel_err(o, oref) < 4e-3
    do = bf(torch.randn(B, Lq, Cn, generator=g)).to(dev)
    dv = torch.empty(B, heads, Lk, d, dtype=torch.bfloat16, device=dev)
    k.gemm(P, do, M=Lk, N=d, K=Lq, a_kmajor=False, lda=ldS, b_kmajor=False, ldb=Cn, out=dv, ldc=d, batch=(B, heads),
           sA=(heads * Lq * ldS, Lq * ldS), sB=(Lq * Cn, d), sC=(heads * Lk * d, Lk * d), backend=backend)
    do4 = do.float().reshape(B, Lq, heads, d).permute(0, 2, 1, 3)
    assert rel_err(dv, P[..., :Lk].float().transpose(-1, -2) @ do4) < 4e-3
    dq = torch.zeros(B, Lq, 3, heads, d, dtype=torch.bfloat16, device=dev)              # dQ written into the qkv-gradient slice
    k.gemm(P, kk, M=Lq, N=d, K=Lk, lda=ldS, b_kmajor=False, ldb=d, out=dq, ldc=3 * Cn, batch=(B, heads),
           sA=(heads * Lq * ldS, Lq * ldS), sB=(heads * Lk * d, Lk * d), sC=(Lq * 3 * Cn, d), backend=backend)
    assert rel_err(dq[:, :, 0].permute(0, 2, 1, 3), P[..., :Lk].float() @ kk.float()) < 4e-3
    assert torch.all(dq[:, :, 1:] == 0)
    if backend == 2 and Lk <= 256:
        # fused epilogues: softmax inside q.k^T, softmax-backward inside dO.v^T
        Pf = torch.full((B, heads, Lq, ldS), float("nan"), dtype=torch.bfloat16, device=dev)
        k.gemm(qkv, kk, M=Lq, N=Lk, K=d, lda=3 * Cn, ldb=d, out=Pf, ldc=ldS, alpha=0.25, act=3, batch=(B, heads),
               sA=(Lq * 3 * Cn, d), sB=(heads * Lk * d, Lk * d), sC=(heads * Lq * ldS, Lq * ldS), backend=backend)
        assert (Pf[..., :Lk].float() - Sref.softmax(-1)).abs().max() < 4e-3
        assert torch.all(Pf[..., Lk:] == 0)
        dSf = torch.full((B, heads, Lq, ldS), float("nan"), dtype=torch.bfloat16, device=dev)
        k.gemm(do, vv, M=Lq, N=Lk, K=d, lda=Cn, ldb=d, out=dSf, ldc=ldS, alpha=0.25, act=4, Z=P, batch=(B, heads),
               sA=(Lq * Cn, d), sB=(heads * Lk * d, Lk * d), sC=(heads * Lq * ldS, Lq * ldS), backend=backend)
        dPref = do4 @ vv.float().transpose(-1, -2)
        Pp = P[..., :Lk].float()
        dSref = 0.25 * Pp * (dPref - (dPref * Pp).sum(-1, keepdim=True))
        assert rel_err(dSf[..., :Lk], dSref) < 8e-3, rel_err(dSf[..., :Lk], dSref)
        assert torch.all(dSf[..., Lk:] == 0)


@pytest.mark.parametrize("backend", [1, 2])
@pytest.mark.parametrize("M,N,Kd,split", [(288, 96, 4096, 8), (96, 384, 8192, 16), (768, 3072, 2048, 1), (192, 192, 1000 * 8, 4),
                                          (256, 768, 8, 1)])
def test_gemm_weight_gradient_layout(backend, M, N, Kd, split):
    """dW[M,N] = dY^T . X with both operands token-major (contraction over rows): MN-major tcgen05
    descriptors / ldmatrix.trans, split-K combined with f32 atomics."""
    k = K()
    g = torch.Generator(device="cpu").manual_seed(M + N + Kd)
    dY = bf(torch.randn(Kd, M, generator=g)).to(dev)
    X = bf(torch.randn(Kd, N, generator=g)).to(dev)
    ref = dY.float().t() @ X.float()
    out = k.gemm(dY, X, M=M, N=N, K=Kd, a_kmajor=False, b_kmajor=False, out_dtype=torch.float32, split_k=split, backend=backend)
    assert rel_err(out, ref) < 2e-5, rel_err(out, ref)
    base = torch.randn(M, N, generator=g).to(dev)
    acc = base.clone()
    k.gemm(dY, X, M=M, N=N, K=Kd, a_kmajor=False, b_kmajor=False, out=acc, accumulate=True, split_k=split, backend=backend)
    assert rel_err(acc, ref + base) < 2e-5
    # fused bias gradient: row sums of dY^T accumulate into `rowsum` from the same kernel (tcgen05: an extra N = 16 MMA
    # against an all-ones tile; mma.sync backend: a column-sum pass), for bf16 and f16 operands
    for cast in (bf, hf):
        dYc, Xc = cast(dY.float()), cast(X.float())
        rs = torch.full((M,), 0.25, device=dev)
        out = k.gemm(dYc, Xc, M=M, N=N, K=Kd, a_kmajor=False, b_kmajor=False, out_dtype=torch.float32, split_k=split, backend=backend,
                     rowsum=rs)
        assert rel_err(out, dYc.float().t() @ Xc.float()) < 2e-5
        assert rel_err(rs, 0.25 + dYc.float().sum(0)) < 2e-5, rel_err(rs, 0.25 + dYc.float().sum(0))


@pytest.mark.parametrize("backend", [1, 2])
@pytest.mark.parametrize("M,N,Kd", [(640, 3072, 768), (1568, 96, 288), (300, 768, 3072), (2080, 192, 768)])
def test_gemm_data_gradient_layout(backend, M, N, Kd):
    """dX = dY . W with the forward's own (N_out, K_in) bf16 weight as an MN-major B operand (no
    transposed copy): plain, times-Z (GELU backward) and bf16-accumulate epilogues."""
    k = K()
    g = torch.Generator(device="cpu").manual_seed(M + 3 * N + Kd)
    dY = bf(torch.randn(M, Kd, generator=g)).to(dev)
    W = bf(torch.randn(Kd, N, generator=g) * 0.05).to(dev)            # Linear(N -> Kd).weight
    ref = dY.float() @ W.float()
    o = k.gemm(dY, W, M=M, N=N, K=Kd, b_kmajor=False, backend=backend)
    assert rel_err(o, ref) < 4e-3
    Z = bf(torch.rand(M, N, generator=g)).to(dev)
    o = k.gemm(dY, W, M=M, N=N, K=Kd, b_kmajor=False, act=2, Z=Z, out_dtype=torch.float32 if backend == 1 else torch.bfloat16,
               backend=backend)
    assert rel_err(o, ref * Z.float()) < 4e-3
    base = bf(torch.randn(M, N, generator=g)).to(dev)
    acc = base.clone()
    k.gemm(dY, W, M=M, N=N, K=Kd, b_kmajor=False, out=acc, accumulate=True, backend=backend)
    assert rel_err(acc, base.float() + ref) < 6e-3


@pytest.mark.parametrize("backend", [1, 2])
def test_gemm_rowscale_and_bf16_accumulate(backend):
    k = K()
    M, N, Kd = 4 * 300, 192, 96
    g = torch.Generator(device="cpu").manual_seed(77)
    A = bf(torch.randn(M, Kd, generator=g)).to(dev)
    B = bf(torch.randn(N, Kd, generator=g) * 0.1).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev)
    rs = torch.tensor([0.0, 1.25, 1.25, 0.0], device=dev)
    ref = (A.float() @ B.float().t() + bias) * rs.repeat_interleave(300)[:, None] + res
    o = k.gemm(A, B, M=M, N=N, K=Kd, bias=bias, residual=res, out_dtype=torch.float32, row_scale=rs, rows_per_scale=300, backend=backend)
    assert rel_err(o, ref) < 2e-5
    assert torch.equal(o[:300], res[:300])
    base = bf(torch.randn(M, N, generator=g)).to(dev)
    acc = base.clone()
    k.gemm(A, B, M=M, N=N, K=Kd, out=acc, accumulate=True, backend=backend)
    assert rel_err(acc, base.float() + A.float() @ B.float().t()) < 6e-3


def test_gemm_ragged_k_and_splitk():
    k = K()
    M, N, Kd = 32, 768, 4096 + 8
    g = torch.Generator(device="cpu").manual_seed(12)
    A = bf(torch.randn(M, Kd, generator=g)).to(dev)
    B = bf(torch.randn(N, Kd, generator=g) * 0.02).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    ref = A.float() @ B.float().t() + bias
    out = k.gemm(A, B, M=M, N=N, K=Kd, bias=bias, out_dtype=torch.float32, split_k=16)
    assert rel_err(out, ref) < 2e-5
    # K not a multiple of 8 with a padded leading dimension (attention P @ V with Nk = 260)
    P = bf(torch.rand(64, 264, generator=g)).to(dev)
    V = bf(torch.randn(260, 96, generator=g)).to(dev)
    o = k.gemm(P, V, M=64, N=96, K=260, lda=264, b_kmajor=False, out_dtype=torch.float32)
    assert rel_err(o, P[:, :260].float() @ V.float()) < 2e-5


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("rows,width", [(1000, 96), (333, 192), (260, 384), (64, 768), (7, 96)])
def test_layernorm_fwd_bwd(rows, width):
    k = K()
    g = torch.Generator(device="cpu").manual_seed(rows)
    x = (torch.randn(rows, width, generator=g) * 2 + 0.5).to(dev)
    gamma = (1 + 0.1 * torch.randn(width, generator=g)).to(dev)
    beta = (0.1 * torch.randn(width, generator=g)).to(dev)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (width,), gr, br, 1e-6)
    y, mean, rstd = k.layernorm_fwd(x, gamma, beta, 1e-6, out_dtype=torch.float32)
    assert rel_err(y, yr) < 1e-5
    y16, _, _ = k.layernorm_fwd(x, gamma, beta, 1e-6)
    assert rel_err(y16, yr) < 4e-3
    dy = bf(torch.randn(rows, width, generator=g)).to(dev)
    add = torch.randn(rows, width, generator=g).to(dev)
    yr.backward(dy.float())
    dg = torch.zeros(width, device=dev)
    db = torch.zeros(width, device=dev)
    dx = k.layernorm_bwd(dy, x, mean, rstd, gamma, dg, db, add=add)
    assert rel_err(dx - add, xr.grad) < 1e-4
    assert rel_err(dg, gr.grad) < 1e-4
    assert rel_err(db, br.grad) < 1e-4
    # second output: the 16-bit, per-row-group scaled copy of dx (operand of the next backward GEMM), same pass
    groups = 4 if rows % 4 == 0 else 1
    rsc = torch.tensor([0.0, 1.25, 1.0, 2.0][:groups], device=dev)
    for dt16 in (torch.bfloat16, torch.float16):
        dg2, db2 = torch.zeros(width, device=dev), torch.zeros(width, device=dev)
        dx2, dx16 = k.layernorm_bwd(dy, x, mean, rstd, gamma, dg2, db2, add=add, copy16=dt16, row_scale=rsc, rows_per_scale=rows // groups)
        assert torch.equal(dx2, dx) and dx16.dtype == dt16
        assert torch.equal(dx16, (dx * rsc.repeat_interleave(rows // groups)[:, None]).to(dt16))


# ------------------------------------------------------------------------------------------------ softmax
@pytest.mark.parametrize("n,ldp", [(64, 64), (256, 256), (260, 264), (1024, 1024), (8, 8)])
def test_softmax_fwd_bwd(n, ldp):
    k = K()
    rows = 3 * n if n == 260 else 512
    g = torch.Generator(device="cpu").manual_seed(n)
    S = torch.full((rows, ldp), float("nan"))
    S[:, :n] = torch.randn(rows, n, generator=g) * 3
    S = S.to(dev)
    mask_hw, mask_t = (64, 4) if n == 260 else (0, 0)
    ref_in = S[:, :n].clone()
    if mask_hw:
        import csts_oracle as O
        ref_in = (ref_in.reshape(3, n, n) - O.spatial_mask((4, 8, 8), dev)).reshape(rows, n)
    Pr = ref_in.softmax(-1)
    P = k.softmax_fwd(S, n, ldp, nq=n, mask_hw=mask_hw, mask_t=mask_t)
    assert P.shape == (rows, ldp)
    assert torch.all(P[:, n:] == 0)
    assert (P[:, :n].float() - Pr).abs().max() < 4e-3
    dP = torch.randn(rows, ldp, generator=g).to(dev)
    dS = k.softmax_bwd(P, dP, n, 0.125)
    Pf = P[:, :n].float()
    ref = 0.125 * Pf * (dP[:, :n] - (dP[:, :n] * Pf).sum(-1, keepdim=True))
    assert rel_err(dS[:, :n], ref) < 5e-3
    assert torch.all(dS[:, n:] == 0)


# ------------------------------------------------------------------------------------------------ dwconv (+LN)
POOL_CASES = [
    # B, heads, d, thw, stride, transposed
    (2, 1, 96, (4, 16, 16), (1, 8, 8), False),
    (2, 2, 96, (4, 16, 16), (1, 2, 2), False),
    (1, 2, 96, (2, 8, 8), (1, 1, 1), False),
    (2, 2, 192, (4, 8, 8), (1, 4, 4), False),
    (2, 2, 96, (4, 4, 4), (1, 2, 2), True),
    (1, 2, 96, (4, 8, 8), (2, 1, 1), True),
    (1, 2, 192, (2, 4, 4), (1, 2, 2), True),
]


@pytest.mark.parametrize("B,heads,d,thw,stride,transposed", POOL_CASES)
def test_dwconv_ln_fwd_bwd(B, heads, d, thw, stride, transposed):
    import csts_oracle as O
    k = K()
    g = torch.Generator(device="cpu").manual_seed(d + thw[1] + stride[1])
    N = thw[0] * thw[1] * thw[2]
    Cn = heads * d
    qkv = bf(torch.randn(B, N, 3, heads, d, generator=g)).to(dev)       # the layout the qkv GEMM writes
    w = (torch.randn(d, 1, 3, 3, 3, generator=g) * 0.2).to(dev)
    gamma = (1 + 0.1 * torch.randn(d, generator=g)).to(dev)
    beta = (0.1 * torch.randn(d, generator=g)).to(dev)
    which = 1
    # oracle (fp32 torch on the same bf16-rounded input)
    t = qkv[:, :, which].float().permute(0, 2, 1, 3).contiguous().requires_grad_(True)   # (B,h,N,d)
    wr, gr, br = w.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr, thw_o = O.pool_tokens(t, thw, wr, stride, gr, br, transposed=transposed)
    in_strides = (N * 3 * Cn, d, 3 * Cn)
    y, pre, mean, rstd, thw_out = k.dwconv(qkv, in_strides, which * Cn, B, heads, d, thw, stride, w, transposed=transposed,
                                           norm=(gamma, beta))
    assert tuple(thw_out) == tuple(thw_o)
    assert rel_err(y, yr) < 8e-3, rel_err(y, yr)
    # backward: LN bwd (generic kernel on the saved pre-LN tensor), then data-grad gather + weight grad
    dy = bf(torch.randn(yr.shape, generator=g)).to(dev)
    yr.backward(dy.float())
    dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    du = k.layernorm_bwd(dy, pre, mean, rstd, gamma, dg, db, dx_dtype=torch.bfloat16)
    assert rel_err(dg, gr.grad) < 2e-2 and rel_err(db, br.grad) < 1e-2
    Lo = thw_out[0] * thw_out[1] * thw_out[2]
    dense = (heads * Lo * d, Lo * d, d)
    dqkv = torch.zeros_like(qkv)
    k.dwconv(du, dense, 0, B, heads, d, thw_out, stride, w, transposed=not transposed,
             out=dqkv, out_strides=in_strides, out_off=which * Cn, thw_out=thw)
    got = dqkv[:, :, which].float().permute(0, 2, 1, 3)
    assert rel_err(got, t.grad) < 1.5e-2, rel_err(got, t.grad)
    assert torch.all(dqkv[:, :, 0] == 0) and torch.all(dqkv[:, :, 2] == 0)
    dw = torch.zeros_like(w)
    if transposed:   # small = conv input, big = d(out)
        k.dwconv_wgrad(qkv, in_strides, which * Cn, thw, du, dense, 0, thw_out, B, heads, d, stride, dw)
    else:            # small = d(out), big = conv input
        k.dwconv_wgrad(du, dense, 0, thw_out, qkv, in_strides, which * Cn, thw, B, heads, d, stride, dw)
    assert rel_err(dw, wr.grad) < 1.5e-2, rel_err(dw, wr.grad)


# ------------------------------------------------------------------------------------------------ skip paths
def test_maxpool_fwd_bwd():
    k = K()
    B, thw, Cn = 2, (2, 8, 12), 96
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(B, thw[0] * thw[1] * thw[2], Cn, generator=g).to(dev)
    xr = x.clone().requires_grad_(True)
    grid = xr.reshape(B, *thw, Cn).permute(0, 4, 1, 2, 3)
    yr = F.max_pool3d(grid, (1, 3, 3), (1, 2, 2), (0, 1, 1)).permute(0, 2, 3, 4, 1).reshape(B, -1, Cn)
    y, arg = k.maxpool_fwd(x, B, thw, Cn)
    assert torch.equal(y, yr)
    dy = torch.randn(y.shape, generator=g).to(dev)
    yr.backward(dy)
    dx = k.maxpool_bwd(dy, arg, B, thw, Cn)
    assert rel_err(dx, xr.grad) < 1e-6


@pytest.mark.parametrize("factors", [(1, 2, 2), (2, 1, 1)])
def test_upsample_fwd_bwd(factors):
    k = K()
    B, thw, Cn = 2, (3, 4, 6), 96
    g = torch.Generator(device="cpu").manual_seed(2)
    x = torch.randn(B, thw[0] * thw[1] * thw[2], Cn, generator=g).to(dev)
    xr = x.clone().requires_grad_(True)
    grid = xr.reshape(B, *thw, Cn).permute(0, 4, 1, 2, 3)
    yr = F.interpolate(grid, scale_factor=tuple(float(f) for f in factors), mode="trilinear").permute(0, 2, 3, 4, 1).reshape(B, -1, Cn)
    y = k.upsample_fwd(x, B, thw, Cn, factors)
    assert rel_err(y, yr) < 1e-6
    dy = torch.randn(y.shape, generator=g).to(dev)
    yr.backward(dy)
    dx = k.upsample_bwd(dy, B, thw, Cn, factors)
    assert rel_err(dx, xr.grad) < 1e-6
    base = torch.randn(dx.shape, generator=g).to(dev)
    dx2 = k.upsample_bwd(dy, B, thw, Cn, factors, dx=base.clone())
    assert rel_err(dx2, xr.grad + base) < 1e-6


# ------------------------------------------------------------------------------------------------ stem / head
@pytest.mark.parametrize("Cin", [3, 1])
def test_patch_embed_im2col_gemm(Cin):
    k = K()
    B, T, H, W = 2, 4, 32, 32
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(B, Cin, T, H, W, generator=g).to(dev)
    w = (torch.randn(96, Cin, 3, 7, 7, generator=g) * 0.05).to(dev)
    b = torch.randn(96, generator=g).to(dev)
    Kp = (Cin * 147 + 7) // 8 * 8
    patches = k.im2col_patch(x, Kp)
    wp = k.cast_bf16(w.reshape(96, -1), ld_out=Kp)
    sp = torch.randn(1, (H // 4) * (W // 4), 96, generator=g).to(dev)
    tp = torch.randn(1, T // 2, 96, generator=g).to(dev)
    pos = k.pos_embed(sp, tp)
    import csts_oracle as O
    assert torch.equal(pos, O.sep_pos_embed(sp, tp)[0])
    M = patches.shape[0]
    tok = k.gemm(patches, wp, M=M, N=96, K=Kp, bias=b, residual=pos, res_mod=pos.shape[0], out_dtype=torch.float32)
    ref = F.conv3d(bf(x).float(), bf(w).float(), b, stride=(2, 4, 4), padding=(1, 3, 3)).flatten(2).transpose(1, 2) + pos
    assert rel_err(tok.reshape(B, -1, 96), ref) < 2e-5
    # position-embedding gradients
    dY = torch.randn(B, T // 2, pos.shape[0] // (T // 2), 96, generator=g).to(dev)
    dsp, dtm = k.pos_embed_bwd(dY, B, T // 2, dY.shape[2], 96)
    assert rel_err(dsp[0], dY.sum((0, 1))) < 1e-5
    assert rel_err(dtm[0], dY.sum((0, 2))) < 1e-5


def test_permute_and_cast_and_colsum():
    k = K()
    g = torch.Generator(device="cpu").manual_seed(4)
    src = torch.randn(5, 70, 33, generator=g).to(dev)
    assert torch.equal(k.permute_021(src, 5, 70, 33, torch.float32), src.transpose(1, 2).contiguous())
    assert torch.equal(k.permute_021(src, 5, 70, 33, torch.bfloat16), bf(src.transpose(1, 2).contiguous()))
    v = torch.randn(1000, 96, generator=g).to(dev)
    assert torch.equal(k.cast_bf16(v), bf(v))
    X = bf(torch.randn(3000, 288, generator=g)).to(dev)
    assert rel_err(k.colsum(X, 3000, 288), X.float().sum(0)) < 1e-5
    Xf = torch.randn(777, 96, generator=g).to(dev)
    assert rel_err(k.colsum(Xf, 777, 96), Xf.sum(0)) < 1e-5
    # wide (several column blocks), strided rows, a width that takes the 4-column kernel, tiny M, f16, accumulate-into
    for M, N, ld, dt_ in [(5000, 3072, 3072, torch.bfloat16), (1234, 768, 2304, torch.bfloat16), (333, 100, 100, torch.float32),
                          (3, 256, 256, torch.float32), (4097, 1536, 1536, torch.float16), (70000, 8, 8, torch.bfloat16)]:
        Xs = torch.randn(M, ld, generator=g).to(dt_).to(dev)
        out = torch.full((N,), 0.5, device=dev)
        k.colsum(Xs, M, N, ld=ld, out=out)
        assert rel_err(out, 0.5 + Xs[:, :N].float().sum(0)) < 2e-5, (M, N, ld, dt_)
    a, b = torch.randn(4096, generator=g).to(dev), torch.randn(4096, generator=g).to(dev)
    assert torch.equal(k.add_f32(a, b), a + b)
    s = torch.tensor([0.25], device=dev)
    assert torch.equal(k.scale_f32(a, s), a * 0.25)


def test_reweight_and_token_mean():
    k = K()
    B, T, S, Cn = 2, 4, 64, 768
    g = torch.Generator(device="cpu").manual_seed(6)
    x = torch.randn(B, T * S, Cn, generator=g).to(dev)
    av = torch.randn(B, 2 * T, Cn, generator=g).to(dev)
    for off in (0, T * Cn):
        w = av.reshape(B, -1)[:, off: off + T * Cn].reshape(B, T, 1, Cn)
        ref = (x.reshape(B, T, S, Cn) * w).reshape(B, T * S, Cn)
        out = k.reweight_fwd(x, av, off, 2 * T * Cn, B, T, S, Cn)
        assert torch.equal(out, ref)
        dout = torch.randn(out.shape, generator=g).to(dev)
        dav = torch.zeros_like(av)
        dx = k.reweight_bwd(dout, x, av, off, 2 * T * Cn, dav, B, T, S, Cn)
        assert rel_err(dx, (dout.reshape(B, T, S, Cn) * w).reshape(B, T * S, Cn)) < 1e-6
        dwr = (dout * x).reshape(B, T, S, Cn).sum(2)
        assert rel_err(dav.reshape(B, -1)[:, off: off + T * Cn].reshape(B, T, Cn), dwr) < 1e-5
    m = k.token_mean_fwd(x, B, T * S, Cn)
    assert rel_err(m, x.mean(1)) < 4e-3
    dm = torch.randn(B, Cn, generator=g).to(dev)
    dx = k.token_mean_bwd(dm, B, T * S, Cn)
    assert rel_err(dx, (dm / (T * S))[:, None, :].expand(B, T * S, Cn)) < 1e-6


def test_classifier_fwd_bwd():
    k = K()
    B, Ti, S, Cn = 2, 4, 64, 96
    g = torch.Generator(device="cpu").manual_seed(8)
    feat = torch.randn(B, 2 * Ti * S, Cn, generator=g).to(dev)
    stem = torch.randn(B, Ti * S, Cn, generator=g).to(dev)
    w = (torch.randn(1, Cn, 1, 1, 1, generator=g) * 0.1).to(dev)
    b = torch.randn(1, generator=g).to(dev)
    fr, sr = feat.clone().requires_grad_(True), stem.clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    f5 = fr.reshape(B, 2 * Ti, 8, 8, Cn).permute(0, 4, 1, 2, 3)
    s5 = sr.reshape(B, Ti, 8, 8, Cn).permute(0, 4, 1, 2, 3)
    ref = F.conv3d(f5 + F.interpolate(s5, size=(2 * Ti, 8, 8), mode="trilinear"), wr, br)
    logits = k.classifier_fwd(feat, stem, w, b, B, Ti, S, Cn)
    assert rel_err(logits.reshape(-1), ref.reshape(-1)) < 1e-5
    dl = torch.randn(ref.shape, generator=g).to(dev)
    ref.backward(dl)
    dfeat, dstem, dw, db = k.classifier_bwd(dl, feat, stem, w, B, Ti, S, Cn)
    assert rel_err(dfeat, fr.grad) < 1e-5
    assert rel_err(dstem, sr.grad) < 1e-5
    assert rel_err(dw, wr.grad.reshape(-1)) < 1e-4
    assert rel_err(db, br.grad) < 1e-4


# ------------------------------------------------------------------------------------------------ losses
def test_kldiv_and_egonce(golden_dir):
    import os
    import csts_oracle as O
    k = K()
    rec = torch.load(os.path.join(golden_dir, "losses.pt"), weights_only=False)
    logits = rec["logits"].to(dev).requires_grad_(True)
    hm = rec["hm"].to(dev)
    loss, prob, dlog = k.kldiv_frame_softmax(logits.detach(), hm, 2.0, T=8)
    assert rel_err(prob, rec["p"].to(dev)) < 1e-5
    assert abs(loss.item() - rec["kld"].item()) < 1e-5 * abs(rec["kld"].item()) + 1e-7
    O.kldiv(O.frame_softmax(logits, 2.0), hm).backward()
    assert rel_err(dlog, logits.grad) < 1e-4
    v = rec["v"].to(dev).requires_grad_(True)
    a = rec["a"].to(dev).requires_grad_(True)
    sim, na, nb = k.sim_matrix_fwd(v.detach(), a.detach())
    assert rel_err(sim, rec["sim"].to(dev)) < 1e-5
    nce, dsim = k.egonce(sim)
    assert abs(nce.item() - rec["nce"].item()) < 1e-4 * abs(rec["nce"].item())
    O.egonce(O.sim_matrix(v, a)).backward()
    dv, da = k.sim_matrix_bwd(v.detach(), a.detach(), sim, dsim, na, nb)
    # logits are sim/0.05 (|z| up to 20): fp32 exp amplifies rounding, hence 1e-3
    assert rel_err(dv, v.grad) < 1e-3
    assert rel_err(da, a.grad) < 1e-3
    # KLDiv with target=None: the uniform-prior (negative entropy) branch, slowfast/models/losses.py:67-71
    from csts_b200.host import losses as L
    from csts_b200.host.utils import frame_softmax
    lg = rec["logits"].to(dev).requires_grad_(True)
    got = L.KLDiv()(frame_softmax(lg, temperature=2))
    got.backward()
    p = O.frame_softmax(logits, 2.0)
    Bn, T, HW = p.shape[0], p.shape[2], p.shape[3] * p.shape[4]
    pm = p.reshape(Bn, T, -1)
    want = (((pm * torch.log(pm + 1e-10)).sum(-1) - math.log(1.0 / HW)).sum(-1) / (T * math.log(HW))).mean()
    logits.grad = None
    want.backward()
    assert abs(got.item() - want.item()) < 1e-5 * abs(want.item()) + 1e-7, (got.item(), want.item())
    assert rel_err(lg.grad, logits.grad) < 1e-4


# ------------------------------------------------------------------------------------------------ storage types


@pytest.mark.parametrize("backend", [1, 2])
def test_gemm_f16_operands(backend):
    """fp16 storage mode: both operands f16 (tcgen05 a_format = b_format = F16; mma.sync .f16.f16)."""
    k = K()
    M, N, Kd = 640, 384, 192
    g = torch.Generator(device="cpu").manual_seed(11)
    A = hf(torch.randn(M, Kd, generator=g)).to(dev)
    B = hf(torch.randn(N, Kd, generator=g) * 0.1).to(dev)
    ref = A.float() @ B.float().t()
    o = k.gemm(A, B, M=M, N=N, K=Kd, out_dtype=torch.float32, backend=backend)
    assert rel_err(o, ref) < 2e-5, rel_err(o, ref)
    # MN-major variants (weight gradient: both operands token-major; data gradient: MN-major weight)
    At, Bt = A.t().contiguous(), B.t().contiguous()
    o = k.gemm(At, Bt, M=M, N=N, K=Kd, a_kmajor=False, b_kmajor=False, out_dtype=torch.float32, backend=backend)
    assert rel_err(o, ref) < 2e-5
    o = k.gemm(A, Bt, M=M, N=N, K=Kd, b_kmajor=False, out_dtype=torch.float32, backend=backend)
    assert rel_err(o, ref) < 2e-5
    o = k.gemm(At, Bt, M=M, N=N, K=Kd, a_kmajor=False, b_kmajor=False, out_dtype=torch.float32, split_k=3, backend=backend)
    assert rel_err(o, ref) < 2e-5
    for odt, otol in ((torch.float16, 6e-4), (torch.bfloat16, 4e-3)):
        o = k.gemm(A, B, M=M, N=N, K=Kd, out_dtype=odt, backend=backend)
        assert o.dtype == odt and rel_err(o, ref) < otol


def test_gemm_mixed_operand_types_take_the_generic_kernel():
    """One tcgen05 kind::f16 MMA cannot mix an f16 and a bf16 operand (it faults on sm_100a): such a product
    is refused by the tcgen05 backend and routed to the mma.sync kernel, which re-rounds the f16 fragment."""
    k = K()
    M, N, Kd = 256, 192, 96
    g = torch.Generator(device="cpu").manual_seed(12)
    A = bf(torch.randn(M, Kd, generator=g)).to(dev)
    B = hf(torch.randn(N, Kd, generator=g) * 0.1).to(dev)
    ref = A.float() @ B.float().t()
    assert rel_err(k.gemm(A, B, M=M, N=N, K=Kd, out_dtype=torch.float32), ref) < 4e-3                  # auto -> mma.sync
    assert rel_err(k.gemm(B, A, M=N, N=M, K=Kd, out_dtype=torch.float32), ref.t()) < 4e-3
    with pytest.raises(RuntimeError, match="unsupported"):
        k.gemm(A, B, M=M, N=N, K=Kd, out_dtype=torch.float32, backend=2)


@pytest.mark.parametrize("backend", [1, 2])
def test_gemm_epilogues_f16_storage(backend):
    """GELU / GELU' / times-Z / accumulate epilogues with f16 operands, C and Z."""
    k = K()
    M, N, Kd = 640, 384, 192
    g = torch.Generator(device="cpu").manual_seed(5)
    A = hf(torch.randn(M, Kd, generator=g)).to(dev)
    B = hf(torch.randn(N, Kd, generator=g) * 0.1).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    pre = A.float() @ B.float().t() + bias
    Z = torch.empty(M, N, dtype=torch.float16, device=dev)
    h = k.gemm(A, B, M=M, N=N, K=Kd, bias=bias, act=1, Z=Z, backend=backend)
    pr = pre.clone().requires_grad_(True)
    (gpre,) = torch.autograd.grad(F.gelu(pr).sum(), pr)
    assert h.dtype == torch.float16 and rel_err(Z, gpre) < 6e-4 and rel_err(h, F.gelu(pre)) < 6e-4
    dY = hf(torch.randn(M, Kd, generator=g)).to(dev)
    W = hf(torch.randn(Kd, N, generator=g) * 0.05).to(dev)
    o = k.gemm(dY, W, M=M, N=N, K=Kd, b_kmajor=False, act=2, Z=Z, out_dtype=torch.float32 if backend == 1 else torch.float16,
               backend=backend)
    assert rel_err(o, (dY.float() @ W.float()) * Z.float()) < 6e-4
    base = hf(torch.randn(M, N, generator=g)).to(dev)
    acc = base.clone()
    k.gemm(A, B, M=M, N=N, K=Kd, out=acc, accumulate=True, backend=backend)
    assert rel_err(acc, base.float() + A.float() @ B.float().t()) < 8e-4


def test_fused_softmax_f16_probabilities():
    k = K()
    B, h, Lq, Lk, d = 2, 2, 256, 200, 96
    ldS = (Lk + 7) // 8 * 8
    g = torch.Generator(device="cpu").manual_seed(3)
    q = hf(torch.randn(B, h, Lq, d, generator=g)).to(dev)
    kk = hf(torch.randn(B, h, Lk, d, generator=g)).to(dev)
    v = hf(torch.randn(B, h, Lk, d, generator=g)).to(dev)
    do = hf(torch.randn(B, h, Lq, d, generator=g)).to(dev)
    scale = d ** -0.5
    P = torch.full((B, h, Lq, ldS), float("nan"), dtype=torch.float16, device=dev)
    k.gemm(q, kk, M=Lq, N=Lk, K=d, out=P, ldc=ldS, alpha=scale, act=3, batch=(B, h), sA=(h * Lq * d, Lq * d), sB=(h * Lk * d, Lk * d),
           sC=(h * Lq * ldS, Lq * ldS))
    Pr = (q.float() @ kk.float().transpose(-1, -2) * scale).softmax(-1)
    assert rel_err(P[..., :Lk], Pr) < 6e-4 and torch.all(P[..., Lk:] == 0)
    dS = torch.full((B, h, Lq, ldS), float("nan"), dtype=torch.float16, device=dev)
    k.gemm(do, v, M=Lq, N=Lk, K=d, out=dS, ldc=ldS, alpha=scale, act=4, Z=P, batch=(B, h), sA=(h * Lq * d, Lq * d),
           sB=(h * Lk * d, Lk * d), sC=(h * Lq * ldS, Lq * ldS))
    dP = do.float() @ v.float().transpose(-1, -2)
    Pf = P[..., :Lk].float()
    dSr = scale * Pf * (dP - (dP * Pf).sum(-1, keepdim=True))
    assert rel_err(dS[..., :Lk], dSr) < 8e-4 and torch.all(dS[..., Lk:] == 0)
    # unfused kernels on the same types
    S = (q.float() @ kk.float().transpose(-1, -2) * scale)
    Sp = torch.zeros(B, h, Lq, ldS, device=dev)
    Sp[..., :Lk] = S
    P2 = k.softmax_fwd(Sp, Lk, ldS, nq=Lq, dtype=torch.float16)
    assert P2.dtype == torch.float16 and rel_err(P2[..., :Lk], Pr) < 6e-4
    dPp = torch.zeros(B, h, Lq, ldS, device=dev)
    dPp[..., :Lk] = dP
    dS2 = k.softmax_bwd(P2, dPp, Lk, scale)
    assert dS2.dtype == torch.float16 and rel_err(dS2[..., :Lk], dSr) < 8e-4
    assert k.softmax_bwd(P2, dPp, Lk, scale, dtype=torch.bfloat16).dtype == torch.bfloat16


def test_rowwise_and_pool_kernels_f16_storage():
    k = K()
    g = torch.Generator(device="cpu").manual_seed(9)
    x = torch.randn(300, 384, generator=g).to(dev)
    gamma, beta = torch.randn(384, generator=g).to(dev), torch.randn(384, generator=g).to(dev)
    y, mean, rstd = k.layernorm_fwd(x, gamma, beta, 1e-6, out_dtype=torch.float16)
    assert y.dtype == torch.float16 and rel_err(y, F.layer_norm(x, (384,), gamma, beta, 1e-6)) < 6e-4
    assert torch.equal(k.cast16(x, torch.float16), hf(x))
    assert torch.equal(k.cast16(x, torch.float16, ld_out=392)[:, :384], hf(x))
    assert torch.equal(k.permute_021(x.view(3, 100, 384), 3, 100, 384, torch.float16), hf(x.view(3, 100, 384).transpose(1, 2).contiguous()))
    # pooling conv + LayerNorm on f16 tokens and its backward pieces with f16 gradients
    B, h, d, thw, stride = 2, 2, 96, (4, 16, 16), (1, 2, 2)
    N = thw[0] * thw[1] * thw[2]
    qkv = hf(torch.randn(B, N, 3, h, d, generator=g)).to(dev)
    w = (torch.randn(d, 1, 3, 3, 3, generator=g) * 0.2).to(dev)
    gq, bq = torch.randn(d, generator=g).to(dev), torch.randn(d, generator=g).to(dev)
    qs = (N * 3 * h * d, d, 3 * h * d)
    out, pre, mean, rstd, thw_o = k.dwconv(qkv, qs, 0, B, h, d, thw, stride, w, norm=(gq, bq), eps=1e-5)
    assert out.dtype == torch.float16 and pre.dtype == torch.float16
    xin = qkv[:, :, 0].float().permute(0, 2, 3, 1).reshape(B * h, d, *thw)             # (B*h, d, T, H, W)
    conv = F.conv3d(xin, w, stride=stride, padding=1, groups=d)
    Lo = thw_o[0] * thw_o[1] * thw_o[2]
    conv_t = conv.reshape(B, h, d, Lo).transpose(2, 3)
    assert rel_err(pre, conv_t) < 6e-4
    assert rel_err(out, F.layer_norm(pre.float(), (d,), gq, bq, 1e-5)) < 8e-4
    du = hf(torch.randn(B, h, Lo, d, generator=g)).to(dev)
    dw = torch.zeros_like(w)
    k.dwconv_wgrad(du, (h * Lo * d, Lo * d, d), 0, thw_o, qkv, qs, 0, thw, B, h, d, stride, dw)
    xr = xin.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    F.conv3d(xr, wr, stride=stride, padding=1, groups=d).backward(du.float().transpose(2, 3).reshape(B * h, d, *thw_o))
    assert rel_err(dw, wr.grad) < 2e-5
    dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    dpre = k.layernorm_bwd(du, pre, mean, rstd, gq, dg, db, dx_dtype=torch.float16)
    pr = pre.float().clone().requires_grad_(True)
    F.layer_norm(pr, (d,), gq, bq, 1e-5).backward(du.float())
    assert dpre.dtype == torch.float16 and rel_err(dpre, pr.grad) < 8e-4
    assert rel_err(dg, (du.float() * F.layer_norm(pre.float(), (d,), None, None, 1e-5)).sum((0, 1, 2))) < 1e-4
    assert rel_err(k.colsum(du.view(-1, d), B * h * Lo, d), du.float().sum((0, 1, 2))) < 1e-5
    # the adjoint (transposed) gather on f16 gradients
    dqkv = torch.zeros(B, N, 3, h, d, dtype=torch.float16, device=dev)
    k.dwconv(du, (h * Lo * d, Lo * d, d), 0, B, h, d, thw_o, stride, w, transposed=True, out=dqkv, out_strides=qs, out_off=0, thw_out=thw)
    assert rel_err(dqkv[:, :, 0].permute(0, 2, 3, 1).reshape(B * h, d, *thw), xr.grad) < 8e-4


@pytest.mark.parametrize("thw,stride,d", [((4, 16, 16), (1, 2, 2), 96), ((2, 8, 8), (1, 1, 1), 96), ((4, 8, 8), (1, 4, 4), 192)])
def test_paired_pool_launch_equals_two_single_launches(thw, stride, d):
    """The k and v pools of a block run as one launch (grid.y = 2): forward (+LayerNorm), adjoint gather and weight
    gradient must equal the single-problem launches bit for bit.  (192-channel heads: the weight gradient runs as two
    96-channel groups per problem, grid.y = 4.)"""
    k = K()
    B, h = 2, 2
    N = thw[0] * thw[1] * thw[2]
    Cn = h * d
    g = torch.Generator(device="cpu").manual_seed(31)
    qkv = bf(torch.randn(B, N, 3, h, d, generator=g)).to(dev)
    wk, wv = ((torch.randn(d, 1, 3, 3, 3, generator=g) * 0.2).to(dev) for _ in range(2))
    nk, nv = ((torch.randn(d, generator=g).to(dev), torch.randn(d, generator=g).to(dev)) for _ in range(2))
    qs = (N * 3 * Cn, d, 3 * Cn)
    rk = k.dwconv(qkv, qs, Cn, B, h, d, thw, stride, wk, norm=nk)
    rv = k.dwconv(qkv, qs, 2 * Cn, B, h, d, thw, stride, wv, norm=nv)
    pk, pv = k.dwconv(qkv, qs, Cn, B, h, d, thw, stride, wk, norm=nk, second=dict(in_off=2 * Cn, w=wv, norm=nv))
    for a, b in zip(rk[:4] + rv[:4], pk[:4] + pv[:4]):
        assert torch.equal(a, b)
    thw_o = rk[4]
    Lo = thw_o[0] * thw_o[1] * thw_o[2]
    dense = (h * Lo * d, Lo * d, d)
    duk, duv = (bf(torch.randn(B, h, Lo, d, generator=g)).to(dev) for _ in range(2))
    one, two = torch.zeros_like(qkv), torch.zeros_like(qkv)
    k.dwconv(duk, dense, 0, B, h, d, thw_o, stride, wk, transposed=True, out=one, out_strides=qs, out_off=Cn, thw_out=thw)
    k.dwconv(duv, dense, 0, B, h, d, thw_o, stride, wv, transposed=True, out=one, out_strides=qs, out_off=2 * Cn, thw_out=thw)
    k.dwconv(duk, dense, 0, B, h, d, thw_o, stride, wk, transposed=True, out=two, out_strides=qs, out_off=Cn, thw_out=thw,
             second=dict(inp=duv, in_off=0, w=wv, out=two, out_off=2 * Cn))
    assert torch.equal(one, two) and one[:, :, 1].abs().sum() > 0 and torch.all(one[:, :, 0] == 0)
    dwk1, dwv1, dwk2, dwv2 = (torch.zeros_like(wk) for _ in range(4))
    k.dwconv_wgrad(duk, dense, 0, thw_o, qkv, qs, Cn, thw, B, h, d, stride, dwk1)
    k.dwconv_wgrad(duv, dense, 0, thw_o, qkv, qs, 2 * Cn, thw, B, h, d, stride, dwv1)
    k.dwconv_wgrad(duk, dense, 0, thw_o, qkv, qs, Cn, thw, B, h, d, stride, dwk2,
                   second=dict(small=duv, small_off=0, big=qkv, big_off=2 * Cn, dw=dwv2))
    assert rel_err(dwk2, dwk1) < 1e-5 and rel_err(dwv2, dwv1) < 1e-5 and dwv1.abs().sum() > 0     # atomics: order differs


# ------------------------------------------------------------------------------------------------ tcgen05 kernel builds
@pytest.mark.parametrize("ctas", [1, 2])
@pytest.mark.parametrize("tile_n", [96, 128, 192])
def test_gemm_tc_builds_and_tile_widths(ctas, tile_n):
    """Both builds of the tcgen05 kernel (one CTA per SM with 12 epilogue warps / two CTAs per SM with 4 epilogue warps and
    256 TMEM columns each) at every tile width they share, on every epilogue: plain, GELU (+GELU'), times-Z, f32 residual,
    accumulate, split-K atomics with the fused row sums, and the MN-major operand layouts."""
    k = K()
    M, N, Kd = 1100, 384, 448                       # ragged M and K tails
    g = torch.Generator(device="cpu").manual_seed(31 + tile_n + ctas)
    A = bf(torch.randn(M, Kd, generator=g)).to(dev)
    B = bf(torch.randn(N, Kd, generator=g) * 0.1).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev)
    kw = dict(backend=2, tile_n=tile_n, ctas=ctas)
    pre = A.float() @ B.float().t() + bias
    Z = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    h = k.gemm(A, B, M=M, N=N, K=Kd, bias=bias, act=1, Z=Z, **kw)
    pr = pre.clone().requires_grad_(True)
    (gpre,) = torch.autograd.grad(F.gelu(pr).sum(), pr)
    assert rel_err(Z, gpre) < 4e-3 and rel_err(h, F.gelu(pre)) < 5e-3
    o = k.gemm(A, B, M=M, N=N, K=Kd, bias=bias, residual=res, out_dtype=torch.float32, **kw)
    assert rel_err(o, pre + res) < 2e-5
    acc = res.clone()
    k.gemm(A, B, M=M, N=N, K=Kd, out=acc, accumulate=True, **kw)
    assert rel_err(acc, res + A.float() @ B.float().t()) < 2e-5
    Bt = B.t().contiguous()                                                   # (K, N): MN-major B (dX = dY . W)
    o = k.gemm(A, Bt, M=M, N=N, K=Kd, b_kmajor=False, act=2, Z=Z, **kw)
    assert rel_err(o, (A.float() @ B.float().t()) * Z.float()) < 4e-3
    # weight-gradient form: both operands token-major, contraction over 1100 tokens, split-K + fused bias gradient
    X = bf(torch.randn(M, 192, generator=g)).to(dev)
    dY = bf(torch.randn(M, N, generator=g)).to(dev)
    want = dY.float().t() @ X.float()
    for split in (1, 3, -1):
        db = torch.zeros(N, device=dev)
        dw = torch.zeros(N, 192, device=dev)
        tn = 192 if tile_n == 128 else tile_n                                # the fused row sums exist for tile widths 96 and 192
        k.gemm(dY, X, M=N, N=192, K=M, a_kmajor=False, b_kmajor=False, lda=N, ldb=192, out=dw, out_is_zero=True, split_k=split,
               rowsum=db, backend=2, tile_n=tn, ctas=ctas)
        assert rel_err(dw, want) < 2e-5, split
        assert rel_err(db, dY.float().sum(0)) < 2e-5, split


def test_gemm_plan_prefers_resident_problems_for_the_two_cta_build():
    """csts_gemm_plan: a problem whose tiles are all resident at once with two CTAs per SM takes that build; an automatic
    split factor keeps >= 4 k-blocks per split and never exceeds the k-block count."""
    import ctypes as C
    from csts_b200 import _lib
    lib = _lib.load()

    def plan(M, N, Kd, a_k=1, b_k=1, split=0, c=1, act=0):
        a = _lib.GemmArgs()
        a.M, a.N, a.K, a.batch1, a.batch2, a.a_kmajor, a.b_kmajor, a.c_dtype, a.act, a.split_k = M, N, Kd, 1, 1, a_k, b_k, c, act, split
        a.a_dtype = a.b_dtype = 1
        bn, ct, sp = C.c_int(), C.c_int(), C.c_int()
        assert lib.csts_gemm_plan(C.byref(a), C.byref(bn), C.byref(ct), C.byref(sp)) == 0
        return bn.value, ct.value, sp.value
    bn, ct, sp = plan(8000, 384, 320)                  # (not a shape of the tuned table: the cost model decides)
    assert ct == 2 and sp == 1 and 384 % bn == 0
    assert plan(8192, 1024, 96, act=3)[:2] == (256, 1)  # whole-row softmax epilogues exist in the one-CTA build only
    bn, ct, sp = plan(1536, 320, 8192, a_k=0, b_k=0, split=-1, c=0)
    assert sp > 1 and (8192 // 64) // sp >= 4
    assert plan(1536, 320, 128, a_k=0, b_k=0, split=-1, c=0)[2] == 1


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("Lq,Lk,heads", [(256, 1024, 2), (1000, 520, 1)])
def test_two_pass_attention_epilogues(Lq, Lk, heads, dtype):
    """Attention with more keys than one accumulator tile holds (attention.py:154-158 at the 1024-key blocks): pass 1 = row
    logsumexp of scale * q.k^T (one CTA walks the n-tiles), pass 2 = P = exp(scale * s - lse) from a second evaluation;
    backward dS = scale * P o (dP - D) with D = rowsum(dO o O) — against torch softmax and its autograd."""
    k = K()
    B, d = 2, 96
    g = torch.Generator(device="cpu").manual_seed(Lq + Lk)
    q = (torch.randn(B, heads, Lq, d, generator=g) * 1.5).to(dtype).to(dev)
    kk = (torch.randn(B, heads, Lk, d, generator=g) * 1.5).to(dtype).to(dev)
    v = torch.randn(B, heads, Lk, d, generator=g).to(dtype).to(dev)
    scale = d ** -0.5
    S = (q.float() @ kk.float().transpose(-1, -2)) * scale
    ldS = (Lk + 7) // 8 * 8
    qk = dict(M=Lq, N=Lk, K=d, lda=d, ldb=d, alpha=scale, batch=(B, heads), sA=(heads * Lq * d, Lq * d), sB=(heads * Lk * d, Lk * d), backend=2)
    lse = torch.empty(B, heads, Lq, device=dev)
    k.gemm(q, kk, out=lse, ldc=Lk, act=5, sC=(heads * Lq, Lq), **qk)
    assert (lse - torch.logsumexp(S, dim=-1)).abs().max() < 2e-4
    P = torch.zeros(B, heads, Lq, ldS, dtype=dtype, device=dev)
    k.gemm(q, kk, out=P, ldc=ldS, act=6, rowvec=lse, sC=(heads * Lq * ldS, Lq * ldS), **qk)
    ref_p = torch.softmax(S, dim=-1)
    tol = 4e-3 if dtype == torch.bfloat16 else 1e-3
    assert rel_err(P[..., :Lk], ref_p) < tol
    assert (P[..., :Lk].float().sum(-1) - 1).abs().max() < 2e-2
    # backward: O = P.V, dO given
    o = (P[..., :Lk].float() @ v.float())                                     # (B, heads, Lq, d)
    o16 = o.permute(0, 2, 1, 3).reshape(B * Lq, heads * d).to(dtype).contiguous()
    do16 = torch.randn(B * Lq, heads * d, generator=g).to(dtype).to(dev)
    D = k.rowdot(do16, o16, B, Lq, heads, d)
    do = do16.float().view(B, Lq, heads, d).permute(0, 2, 1, 3)
    assert rel_err(D, (do * o16.float().view(B, Lq, heads, d).permute(0, 2, 1, 3)).sum(-1)) < 1e-5
    dS = torch.zeros(B, heads, Lq, ldS, dtype=dtype, device=dev)
    Cn = heads * d
    k.gemm(do16, v, M=Lq, N=Lk, K=d, lda=Cn, ldb=d, out=dS, ldc=ldS, alpha=scale, act=7, Z=P, rowvec=D, batch=(B, heads),
           sA=(Lq * Cn, d), sB=(heads * Lk * d, Lk * d), sC=(heads * Lq * ldS, Lq * ldS), backend=2)
    Pf = P[..., :Lk].float()
    dP = do @ v.float().transpose(-1, -2)
    want = scale * Pf * (dP - (dP * Pf).sum(-1, keepdim=True))
    assert rel_err(dS[..., :Lk], want) < (2e-2 if dtype == torch.bfloat16 else 5e-3)
