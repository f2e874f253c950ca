"""Tensor-level wrappers over the C-ABI (one Python function per exported kernel family).

They allocate outputs with torch (the caching allocator owns all memory), compute strides and
call into libcsts_b200.so through ``_lib.call``.  No math happens here.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import BF16, F16, F32, GemmArgs, PoolArgs, WgradArgs, call, dt, ptr

_FORCE_BACKEND = 0   # 0 auto, 1 mma.sync, 2 tcgen05 (tests flip this to cross-check the two GEMMs)


GEMM_PROFILE = None   # when a list: every gemm() appends (start_event, end_event, flops, bytes, used_tcgen05)
GEMM_RECORD = None    # when a list: every gemm() appends (argument block, tensors kept alive, flops, used_tcgen05) — bench.py replays them


def set_gemm_backend(code: int):
    global _FORCE_BACKEND
    _FORCE_BACKEND = int(code)


def build_gemm_args(A, B, *, M, N, K, a_kmajor=True, b_kmajor=True, lda=None, ldb=None, out=None, out_dtype=None,
                    ldc=None, bias=None, act=0, Z=None, residual=None, res_mod=0, accumulate=False, alpha=1.0, split_k=1,
                    batch=(1, 1), sA=(0, 0), sB=(0, 0), sC=(0, 0), a_off=0, b_off=0, c_off=0, backend=None,
                    row_scale=None, rows_per_scale=0, rowsum=None, tile_n=0, ctas=0, out_is_zero=False, want_out=False,
                    rowvec=None):
    """The csts_gemm_args block of a gemm() call (same keywords)."""
    assert A.dtype in (torch.bfloat16, torch.float16) and B.dtype in (torch.bfloat16, torch.float16)
    if out_dtype is None:
        out_dtype = A.dtype
    if lda is None:
        lda = K if a_kmajor else M
    if ldb is None:
        ldb = K if b_kmajor else N
    if out is None:
        assert batch == (1, 1)
        out = torch.empty((M, N), dtype=out_dtype, device=A.device)
    if ldc is None:
        ldc = N
    a = GemmArgs()
    esz = 2
    a.A = A.data_ptr() + a_off * esz
    a.B = B.data_ptr() + b_off * esz
    a.C = out.data_ptr() + c_off * out.element_size()
    a.Z = Z.data_ptr() if Z is not None else None
    a.bias = bias.data_ptr() if bias is not None else None
    a.residual = residual.data_ptr() if residual is not None else None
    a.row_scale = row_scale.data_ptr() if row_scale is not None else None
    a.rowsum = rowsum.data_ptr() if rowsum is not None else None      # f32 [M], += row sums of op(A) (bias gradient of a wgrad)
    a.rowvec = rowvec.data_ptr() if rowvec is not None else None      # f32 [batch][M] (two-pass attention epilogues)
    a.rows_per_scale = rows_per_scale
    a.lda, a.ldb, a.ldc = lda, ldb, ldc
    a.ldz = ldc
    a.ldr = N
    a.sA1, a.sA2 = sA
    a.sB1, a.sB2 = sB
    a.sC1, a.sC2 = sC
    a.M, a.N, a.K = M, N, K
    a.batch1, a.batch2 = batch
    a.a_kmajor, a.b_kmajor = int(a_kmajor), int(b_kmajor)
    a.c_dtype = dt(out)
    a.a_dtype, a.b_dtype = dt(A), dt(B)
    a.z_dtype = dt(Z) if Z is not None else 0
    a.act = act
    a.accumulate = int(accumulate)
    a.res_mod = res_mod
    a.split_k = split_k
    a.alpha = alpha
    a.tile_n, a.ctas = tile_n, ctas
    a.backend = _FORCE_BACKEND if backend is None else backend
    if bias is not None:
        assert bias.dtype == torch.float32
    if residual is not None:
        assert residual.dtype == torch.float32
    if out_is_zero and not accumulate:
        # resolve the split factor first: more than one split accumulates into the zeroed output, a single one overwrites it
        a.accumulate = 1
        if _lib.load().csts_gemm_backend(C.byref(a)) == 2:
            bn_, ct_, sp_ = C.c_int(), C.c_int(), C.c_int()
            _lib.load().csts_gemm_plan(C.byref(a), C.byref(bn_), C.byref(ct_), C.byref(sp_))
            a.split_k, a.tile_n, a.ctas = sp_.value, bn_.value, ct_.value
            a.accumulate = int(sp_.value > 1)
        else:
            a.split_k = max(1, split_k)
            a.accumulate = int(a.split_k > 1)
    return (a, out) if want_out else a


def gemm(A, B, **kw):
    """C = epi(alpha * op(A) @ op(B)).  A/B are bf16 or f16 storage tensors (independently); offsets/strides
    in elements.  out_dtype defaults to A's type.  split_k < 0: the library picks the factor.  out_is_zero: `out`
    holds zeros (a slice of the gradient arena), so a split-K product may accumulate into it without its own memset.
    tile_n / ctas: tuning overrides of the tcgen05 launcher (0 = automatic).  Keywords: see build_gemm_args."""
    a, out = build_gemm_args(A, B, want_out=True, **kw)
    if GEMM_RECORD is not None:
        GEMM_RECORD.append((a, (A, B, out) + tuple(v for v in kw.values() if torch.is_tensor(v)), 2.0 * a.M * a.N * a.K * a.batch1 * a.batch2,
                            _lib.load().csts_gemm_backend(C.byref(a)) == 2))
    if GEMM_PROFILE is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nb = a.batch1 * a.batch2
        ev0.record()
        call("csts_gemm", C.byref(a))
        ev1.record()
        tc = _lib.load().csts_gemm_backend(C.byref(a)) == 2
        GEMM_PROFILE.append((ev0, ev1, 2.0 * a.M * a.N * a.K * nb, nb * (2.0 * (a.M * a.K + a.N * a.K) + out.element_size() * a.M * a.N), tc,
                             (a.M, a.N, a.K, nb, a.a_kmajor, a.b_kmajor, a.act, int(a.residual is not None), a.c_dtype, a.split_k)))
        return out
    call("csts_gemm", C.byref(a))
    return out


def layernorm_fwd(x, gamma, beta, eps, out_dtype=torch.bfloat16, want_stats=True):
    width = x.shape[-1]
    rows = x.numel() // width
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device) if want_stats else None
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if want_stats else None
    call("csts_layernorm_fwd", ptr(x), dt(x), ptr(y), dt(y), ptr(gamma), ptr(beta), ptr(mean), ptr(rstd), rows, width, eps)
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, add=None, dx_dtype=torch.float32, copy16=None, row_scale=None,
                  rows_per_scale=0):
    """dgamma/dbeta (f32, zero-initialised by the caller) are accumulated into.  copy16 (a 16-bit dtype): also
    return a 16-bit copy of dx whose row m is scaled by row_scale[m // rows_per_scale] (written in the same pass)."""
    width = x.shape[-1]
    rows = x.numel() // width
    dx = torch.empty(x.shape, dtype=dx_dtype, device=x.device)
    dx16 = torch.empty(x.shape, dtype=copy16, device=x.device) if copy16 is not None else None
    call("csts_layernorm_bwd", ptr(dy), dt(dy), ptr(x), dt(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(add), ptr(dx), dt(dx),
         ptr(dgamma), ptr(dbeta), rows, width, ptr(dx16), dt(dx16) if dx16 is not None else 0, ptr(row_scale), rows_per_scale)
    return dx if copy16 is None else (dx, dx16)


def layernorm_bwd_pair(dy, x, mean, rstd, gamma, dgamma, dbeta, dx_dtype):
    """Two LayerNorm backward problems of identical shape and types in one launch; every argument is a pair."""
    width = x[0].shape[-1]
    rows = x[0].numel() // width
    assert x[1].shape == x[0].shape and dy[0].dtype == dy[1].dtype and x[0].dtype == x[1].dtype
    dx = (torch.empty(x[0].shape, dtype=dx_dtype, device=x[0].device), torch.empty(x[1].shape, dtype=dx_dtype, device=x[1].device))
    call("csts_layernorm_bwd_pair", ptr(dy[0]), ptr(dy[1]), dt(dy[0]), ptr(x[0]), ptr(x[1]), dt(x[0]), ptr(mean[0]), ptr(mean[1]),
         ptr(rstd[0]), ptr(rstd[1]), ptr(gamma[0]), ptr(gamma[1]), ptr(dx[0]), ptr(dx[1]), dt(dx[0]), ptr(dgamma[0]), ptr(dgamma[1]),
         ptr(dbeta[0]), ptr(dbeta[1]), rows, width)
    return dx


def softmax_fwd(S, n, ldp, nq, mask_hw=0, mask_t=0, dtype=torch.bfloat16):
    """S f32 (..., lds) -> P (bf16 / f16) (..., ldp) with zeroed pad columns."""
    lds = S.shape[-1]
    rows = S.numel() // lds
    P = torch.empty(S.shape[:-1] + (ldp,), dtype=dtype, device=S.device)
    call("csts_softmax_fwd", ptr(S), ptr(P), dt(P), rows, n, lds, ldp, nq, mask_hw, mask_t)
    return P


def softmax_bwd(P, dP, n, scale, dtype=None):
    """dS (16-bit, P's type unless given) from the stored probabilities P (bf16 / f16) and dP (f32)."""
    ldp = P.shape[-1]
    rows = P.numel() // ldp
    dS = torch.empty(P.shape, dtype=P.dtype if dtype is None else dtype, device=P.device)
    call("csts_softmax_bwd", ptr(P), dt(P), ptr(dP), ptr(dS), dt(dS), rows, n, ldp, dP.shape[-1], scale)
    return dS


def rowdot(dO, O, B, Lq, heads, d):
    """D[b, head, q] = sum_d dO[b, q, head, d] * O[b, q, head, d]  (f32) for 16-bit (B*Lq, heads*d) tensors."""
    assert dO.dtype == O.dtype and dO.is_contiguous() and O.is_contiguous()
    D = torch.empty((B, heads, Lq), dtype=torch.float32, device=O.device)
    call("csts_rowdot", ptr(dO), ptr(O), dt(O), ptr(D), B, Lq, heads, d)
    return D


def cast16(src, dtype, ld_out=None, row_scale=None, rows_per_scale=0, out=None):
    """f32 (rows, cols) -> bf16 / f16 (rows, ld_out), zero padded; optional per-row-group scale
    (row m is multiplied by row_scale[m // rows_per_scale]).  `out`: an existing destination to overwrite."""
    src2 = src.reshape(-1, src.shape[-1]) if src.dim() > 1 else src.reshape(1, -1)
    rows, cols = src2.shape
    ld = cols if ld_out is None else ld_out
    if out is not None and (out.dtype != dtype or out.numel() != rows * ld or not out.is_contiguous()):
        out = None
    dst = out.view(rows, ld) if out is not None else torch.empty((rows, ld), dtype=dtype, device=src.device)
    call("csts_cast16", ptr(src2), ptr(dst), dt(dst), rows, cols, ld, ptr(row_scale), rows_per_scale)
    if out is not None:
        return out
    return dst if ld_out is not None else dst.reshape(src.shape)


def cast_bf16(src, ld_out=None, row_scale=None, rows_per_scale=0):
    return cast16(src, torch.bfloat16, ld_out, row_scale, rows_per_scale)


def permute_021(src, a, b, c, out_dtype):
    """f32 [a][b][c] -> [a][c][b] in out_dtype."""
    dst = torch.empty((a, c, b), dtype=out_dtype, device=src.device)
    call("csts_permute_021", ptr(src), ptr(dst), dt(dst), a, b, c)
    return dst


def add_f32(a, b):
    out = torch.empty_like(a)
    call("csts_add_f32", ptr(a), ptr(b), ptr(out), a.numel())
    return out


def scale_f32(a, device_scalar):
    out = torch.empty_like(a)
    call("csts_scale_f32", ptr(a), ptr(device_scalar), ptr(out), a.numel())
    return out


def colsum(X, M, N, ld=None, out=None):
    if out is None:
        out = torch.zeros(N, dtype=torch.float32, device=X.device)
    call("csts_colsum", ptr(X), dt(X), ptr(out), M, N, N if ld is None else ld)
    return out


def _out_grid(thw, stride, transposed):
    if transposed:
        return tuple(n * s for n, s in zip(thw, stride))
    return tuple((n - 1) // s + 1 for n, s in zip(thw, stride))


def dwconv(inp, in_strides, in_off, B, heads, d, thw_in, stride, w, *, transposed=False, norm=None, eps=1e-5,
           out=None, out_strides=None, out_off=0, thw_out=None, second=None):
    """Depthwise 3x3x3 (transposed) conv over a token grid; optional fused LayerNorm(d).

    inp: bf16 / f16 storage (out and pre get the same type); element (b, head, pos, c) at in_off + b*sB + head*sH + pos*sP + c.
    Returns (out, pre, mean, rstd, thw_out); out is dense (B, heads, Lo, d) unless `out` is given.

    second = dict(inp=, in_off=, w=, norm=, out=, out_off=): a second problem of identical geometry and strides (the k
    and v pools of a block) executed by the same launch; the call then returns (first results, second results).
    """
    if thw_out is None:
        thw_out = _out_grid(thw_in, stride, transposed)
    Lo = thw_out[0] * thw_out[1] * thw_out[2]
    dense = out is None
    a = PoolArgs()
    a.dtype = dt(inp)

    def one(inp_, in_off_, w_, norm_, out_, out_off_):
        if dense:
            out_ = torch.empty((B, heads, Lo, d), dtype=inp_.dtype, device=inp_.device)
        assert out_.dtype == inp_.dtype == inp.dtype
        pre = mean = rstd = None
        gamma = beta = None
        if norm_ is not None:
            gamma, beta = norm_
            pre = torch.empty((B, heads, Lo, d), dtype=inp_.dtype, device=inp_.device)
            mean = torch.empty(B * heads * Lo, dtype=torch.float32, device=inp_.device)
            rstd = torch.empty_like(mean)
        ptrs = (inp_.data_ptr() + in_off_ * 2, out_.data_ptr() + out_off_ * 2, w_.data_ptr(), ptr(gamma), ptr(beta), ptr(pre), ptr(mean), ptr(rstd))
        return ptrs, (out_, pre, mean, rstd, thw_out)

    p1, r1 = one(inp, in_off, w, norm, out, out_off)
    a.inp, a.out, a.w, a.gamma, a.beta, a.pre, a.mean, a.rstd = p1
    r2 = None
    if second is not None:
        assert (second.get("norm") is None) == (norm is None)
        p2, r2 = one(second.get("inp", inp), second["in_off"], second["w"], second.get("norm"), second.get("out"), second.get("out_off", 0))
        a.in2, a.out2, a.w2, a.gamma2, a.beta2, a.pre2, a.mean2, a.rstd2 = p2
    if dense:
        out_strides = (heads * Lo * d, Lo * d, d)
    a.in_sB, a.in_sH, a.in_sP = in_strides
    a.out_sB, a.out_sH, a.out_sP = out_strides
    a.B, a.heads, a.d = B, heads, d
    a.Ti, a.Hi, a.Wi = thw_in
    a.To, a.Ho, a.Wo = thw_out
    a.st, a.sh, a.sw = stride
    a.transposed = int(transposed)
    a.eps = eps
    call("csts_dwconv", C.byref(a))
    return r1 if second is None else (r1, r2)


def dwconv_wgrad(small, small_strides, small_off, thw_small, big, big_strides, big_off, thw_big, B, heads, d, stride, dw, second=None):
    """second = dict(small=, small_off=, big=, big_off=, dw=): a second problem of identical geometry in the same launch."""
    a = WgradArgs()
    if second is not None:
        s2, b2 = second["small"], second["big"]
        assert s2.dtype == small.dtype and b2.dtype == big.dtype
        a.small2 = s2.data_ptr() + second["small_off"] * 2
        a.big2 = b2.data_ptr() + second["big_off"] * 2
        a.dw2 = second["dw"].data_ptr()
    a.small = small.data_ptr() + small_off * 2
    a.big = big.data_ptr() + big_off * 2
    a.dw = dw.data_ptr()
    a.small_dtype, a.big_dtype = dt(small), dt(big)
    a.small_sB, a.small_sH, a.small_sP = small_strides
    a.big_sB, a.big_sH, a.big_sP = big_strides
    a.B, a.heads, a.d = B, heads, d
    a.Ts, a.Hs, a.Ws = thw_small
    a.Tb, a.Hb, a.Wb = thw_big
    a.st, a.sh, a.sw = stride
    call("csts_dwconv_wgrad", C.byref(a))
    return dw


def maxpool_fwd(x, B, thw, Cn, want_arg=True):
    T, H, W = thw
    y = torch.empty((B, T * (H // 2) * (W // 2), Cn), dtype=torch.float32, device=x.device)
    arg = torch.empty(y.shape, dtype=torch.uint8, device=x.device) if want_arg else None
    call("csts_maxpool_fwd", ptr(x), ptr(y), ptr(arg), B, T, H, W, Cn)
    return y, arg


def maxpool_bwd(dy, arg, B, thw, Cn):
    T, H, W = thw
    dx = torch.empty((B, T * H * W, Cn), dtype=torch.float32, device=dy.device)
    call("csts_maxpool_bwd", ptr(dy), ptr(arg), ptr(dx), B, T, H, W, Cn)
    return dx


def upsample_fwd(x, B, thw, Cn, factors):
    T, H, W = thw
    ft, fh, fw = factors
    y = torch.empty((B, T * ft * H * fh * W * fw, Cn), dtype=torch.float32, device=x.device)
    call("csts_upsample_fwd", ptr(x), ptr(y), B, T, H, W, Cn, ft, fh, fw)
    return y


def upsample_bwd(dy, B, thw, Cn, factors, dx=None):
    T, H, W = thw
    ft, fh, fw = factors
    acc = dx is not None
    if dx is None:
        dx = torch.empty((B, T * H * W, Cn), dtype=torch.float32, device=dy.device)
    call("csts_upsample_bwd", ptr(dy), ptr(dx), B, T, H, W, Cn, ft, fh, fw, int(acc))
    return dx


def im2col_patch(x, Kp, dtype=torch.bfloat16):
    B, Cin, T, H, W = x.shape
    out = torch.empty((B * (T // 2) * (H // 4) * (W // 4), Kp), dtype=dtype, device=x.device)
    call("csts_im2col_patch", ptr(x), ptr(out), dt(out), B, Cin, T, H, W, Kp)
    return out


def pos_embed(spatial, temporal):
    T, HW, Cn = temporal.shape[-2], spatial.shape[-2], spatial.shape[-1]
    pos = torch.empty((T * HW, Cn), dtype=torch.float32, device=spatial.device)
    call("csts_pos_embed", ptr(spatial), ptr(temporal), ptr(pos), T, HW, Cn)
    return pos


def pos_embed_bwd(dY, B, T, HW, Cn):
    dsp = torch.empty((1, HW, Cn), dtype=torch.float32, device=dY.device)
    dtm = torch.zeros((1, T, Cn), dtype=torch.float32, device=dY.device)
    call("csts_pos_embed_bwd", ptr(dY), ptr(dsp), ptr(dtm), B, T, HW, Cn)
    return dsp, dtm


def reweight_fwd(x, w, w_off, w_sB, B, T, S, Cn):
    out = torch.empty((B, T * S, Cn), dtype=torch.float32, device=x.device)
    call("csts_reweight_fwd", ptr(x), C.c_void_p(w.data_ptr() + 4 * w_off), ptr(out), B, T, S, Cn, w_sB)
    return out


def reweight_bwd(dout, x, w, w_off, w_sB, dw, B, T, S, Cn, want_dx=True):
    dx = torch.empty((B, T * S, Cn), dtype=torch.float32, device=x.device) if want_dx else None
    call("csts_reweight_bwd", ptr(dout), ptr(x), C.c_void_p(w.data_ptr() + 4 * w_off), ptr(dx),
         C.c_void_p(dw.data_ptr() + 4 * w_off), B, T, S, Cn, w_sB)
    return dx


def token_mean_fwd(x, B, N, Cn, dtype=torch.bfloat16):
    out = torch.empty((B, Cn), dtype=dtype, device=x.device)
    call("csts_token_mean_fwd", ptr(x), ptr(out), dt(out), B, N, Cn)
    return out


def token_mean_bwd(dout, B, N, Cn, dx=None):
    acc = dx is not None
    if dx is None:
        dx = torch.empty((B, N, Cn), dtype=torch.float32, device=dout.device)
    call("csts_token_mean_bwd", ptr(dout), ptr(dx), B, N, Cn, int(acc))
    return dx


def classifier_fwd(feat, stem, w, bias, B, Ti, S, Cn):
    logits = torch.empty((B, 1, 2 * Ti, S), dtype=torch.float32, device=feat.device)
    call("csts_classifier_fwd", ptr(feat), ptr(stem), ptr(w), ptr(bias), ptr(logits), B, Ti, S, Cn)
    return logits


def classifier_bwd(dlogits, feat, stem, w, B, Ti, S, Cn):
    dfeat = torch.empty_like(feat)
    dstem = torch.empty_like(stem)
    dw = torch.zeros(Cn, dtype=torch.float32, device=feat.device)
    db = torch.zeros(1, dtype=torch.float32, device=feat.device)
    call("csts_classifier_bwd", ptr(dlogits), ptr(feat), ptr(stem), ptr(w), ptr(dfeat), ptr(dstem), ptr(dw), ptr(db), B, Ti, S, Cn)
    return dfeat, dstem, dw, db


def kldiv_frame_softmax(logits, target, temperature, T, want_grad=True):
    """logits (B,1,T,H,W) f32, target (B,T,H,W) f32 -> (loss[1], prob like logits, dlogits or None)."""
    HW = logits.shape[-1] * logits.shape[-2]
    frames = logits.numel() // HW
    prob = torch.empty_like(logits)
    frame_kl = torch.empty(frames, dtype=torch.float32, device=logits.device)
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    dlogits = torch.empty_like(logits) if want_grad else None
    call("csts_kldiv_frame_softmax", ptr(logits), ptr(target), ptr(prob), ptr(frame_kl), ptr(loss), ptr(dlogits), frames, HW, T,
         temperature)
    return loss, prob, dlogits


def sim_matrix_fwd(a, b, eps=1e-8):
    n, D = a.shape
    sim = torch.empty((n, n), dtype=torch.float32, device=a.device)
    na = torch.empty(n, dtype=torch.float32, device=a.device)
    nb = torch.empty_like(na)
    call("csts_sim_matrix_fwd", ptr(a), ptr(b), ptr(sim), ptr(na), ptr(nb), n, D, eps)
    return sim, na, nb


def sim_matrix_bwd(a, b, sim, dsim, na, nb):
    n, D = a.shape
    da, db = torch.empty_like(a), torch.empty_like(b)
    call("csts_sim_matrix_bwd", ptr(a), ptr(b), ptr(sim), ptr(dsim), ptr(na), ptr(nb), ptr(da), ptr(db), n, D)
    return da, db


def egonce(sim, temperature=0.05, want_grad=True):
    n = sim.shape[0]
    loss = torch.empty(1, dtype=torch.float32, device=sim.device)
    dsim = torch.empty_like(sim) if want_grad else None
    scratch = torch.empty(2 * n, dtype=torch.float32, device=sim.device)
    call("csts_egonce", ptr(sim), ptr(loss), ptr(dsim), ptr(scratch), n, temperature)
    return loss, dsim
