"""One CSTS transformer block (encoder / decoder / spatial-fusion / temporal-fusion) as a single
autograd node running on the libcsts_b200 kernels.

Forward follows ``MultiScaleBlock.forward`` / ``MultiScaleAttention.forward``
(attention.py:238-248, :120-162), the decoder variants (:365-392, :469-479) and the fusion blocks
(av_attention.py:120-152, :229-250, :322-372, :450-473).  Data layout:

  residual stream x            f32  (B, N, C)           token-major
  LayerNorm outputs, qkv, MLP  16b  (B*N, ...)          GEMM operands
  qkv                          16b  (B, N, 3, heads, d) exactly as the qkv GEMM writes it
  pooled q / k / v             16b  (B, heads, L', d)
  attention out                16b  (B, Lq, heads*d)    written strided by the P.V GEMM

"16b" is the storage type of the model's precision mode (weights.Precision): bf16 for activations and
their gradients by default, fp16 for both under TRAIN.MIXED_PRECISION (the reference's fp16 autocast +
GradScaler contract, tools/train_avgaze_net.py:70,99-109).  One tcgen05 MMA cannot mix the two formats,
so a mode uses a single 16-bit type throughout.

The reference's reshape/permute/contiguous copies around the pooling convs and the head
split/merge do not exist here: the pooling and attention kernels take element strides.
"""
import torch

from .. import kernels as K

EPS_BLOCK = 1e-6     # norm1 / norm2: partial(nn.LayerNorm, eps=1e-6)  (custom_multimodal_builder.py:61)
EPS_POOL = 1e-5      # norm_q/k/v: nn.LayerNorm default               (attention.py:206)


def block_param_names(spec):
    """Parameter names of one block in the reference's registration order."""
    n = ["norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias"]
    if spec.stride_q is not None:
        n += [("attn.upsample_q.weight" if spec.kind == "dec" else "attn.pool_q.weight"), "attn.norm_q.weight", "attn.norm_q.bias"]
    if spec.stride_kv is not None:
        n += ["attn.pool_k.weight", "attn.norm_k.weight", "attn.norm_k.bias",
              "attn.pool_v.weight", "attn.norm_v.weight", "attn.norm_v.bias"]
    n += ["norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias"]
    if spec.dim != spec.dim_out:
        n += ["proj.weight", "proj.bias"]
    return n


class _Ref:
    """A (B, heads, L, d) 16-bit tensor addressed by element strides inside `buf`."""
    __slots__ = ("buf", "off", "sB", "sH", "sP", "L", "thw")

    def __init__(self, buf, off, sB, sH, sP, L, thw):
        self.buf, self.off, self.sB, self.sH, self.sP, self.L, self.thw = buf, off, sB, sH, sP, L, thw

    @property
    def strides(self):
        return (self.sB, self.sH, self.sP)


def _dense_ref(t, B, h, L, d, thw):
    return _Ref(t, 0, h * L * d, L * d, d, L, thw)


def block_forward(spec, p, wc, x, thw, dp_scale=None, save=True, want_attn=False):
    """x: f32 (B, N, C).  p: {name: f32 parameter}.  wc: WeightCache.  Returns (y, thw_q, saved, attn).
    dp_scale: None or f32 (2, B) — the per-sample DropPath scales (0 or 1/keep) of the attention branch (row 0)
    and of the MLP branch (row 1): the reference draws them independently (attention.py:242 and :247)."""
    B, N, C = x.shape
    dp_a, dp_m = (dp_scale[0], dp_scale[1]) if dp_scale is not None else (None, None)
    h, d = spec.heads, spec.head_dim
    M = B * N
    dec = spec.kind == "dec"
    xn1, mean1, rstd1 = K.layernorm_fwd(x, p["norm1.weight"], p["norm1.bias"], EPS_BLOCK, out_dtype=wc.act)
    qkv = K.gemm(xn1.view(M, C), wc.w(p["attn.qkv.weight"]), M=M, N=3 * C, K=C, bias=p["attn.qkv.bias"])
    qs = (N * 3 * C, d, 3 * C)                       # (batch, head, position) strides inside qkv
    sv = {}
    # ---- q / k / v, pooled where the block says so -------------------------------------------
    if spec.stride_q is not None:
        wq = p["attn.upsample_q.weight" if dec else "attn.pool_q.weight"]
        qt, q_pre, q_mean, q_rstd, thw_q = K.dwconv(qkv, qs, 0, B, h, d, thw, spec.stride_q, wq, transposed=dec,
                                                    norm=(p["attn.norm_q.weight"], p["attn.norm_q.bias"]), eps=EPS_POOL)
        Lq = thw_q[0] * thw_q[1] * thw_q[2]
        q = _dense_ref(qt, B, h, Lq, d, thw_q)
        sv["q_pool"] = (q_pre, q_mean, q_rstd)
    else:
        thw_q, Lq = tuple(thw), N
        q = _Ref(qkv, 0, *qs, N, thw_q)
    if spec.stride_kv is not None:
        # pool_k and pool_v share their geometry: one launch (grid.y = 2)
        (kt, k_pre, k_mean, k_rstd, thw_k), (vt, v_pre, v_mean, v_rstd, _) = K.dwconv(
            qkv, qs, C, B, h, d, thw, spec.stride_kv, p["attn.pool_k.weight"], norm=(p["attn.norm_k.weight"], p["attn.norm_k.bias"]),
            eps=EPS_POOL, second=dict(in_off=2 * C, w=p["attn.pool_v.weight"], norm=(p["attn.norm_v.weight"], p["attn.norm_v.bias"])))
        Lk = thw_k[0] * thw_k[1] * thw_k[2]
        k = _dense_ref(kt, B, h, Lk, d, thw_k)
        v = _dense_ref(vt, B, h, Lk, d, thw_k)
        sv["k_pool"] = (k_pre, k_mean, k_rstd)
        sv["v_pool"] = (v_pre, v_mean, v_rstd)
    else:
        Lk = N
        k = _Ref(qkv, C, *qs, N, tuple(thw))
        v = _Ref(qkv, 2 * C, *qs, N, tuple(thw))
    # ---- attention: S = scale * q k^T -> softmax (+ in-frame mask) -> P v ---------------------
    ldS = (Lk + 7) // 8 * 8
    scale = d ** -0.5
    # Lk <= 256 (20 of the 26 blocks): softmax runs in the epilogue of the q.k^T kernel, S never reaches HBM
    fused_softmax = Lk <= 256 and Lq >= 64 and spec.kind != "spatial"
    # Lk > 256 (6 blocks, 1024 keys): two passes over q.k^T — row logsumexp first, then P = exp(scale * s - lse) from the
    # epilogue of a second evaluation (the contraction is only d deep); the f32 scores never reach HBM either
    two_pass = not fused_softmax and spec.kind != "spatial" and Lq >= 64 and Lk % 8 == 0 and not want_attn
    qk = dict(M=Lq, N=Lk, K=d, lda=q.sP, ldb=k.sP, alpha=scale, batch=(B, h), sA=(q.sB, q.sH), sB=(k.sB, k.sH), a_off=q.off, b_off=k.off)
    if fused_softmax:
        P = torch.empty((B, h, Lq, ldS), dtype=wc.act, device=x.device)
        K.gemm(q.buf, k.buf, out=P, ldc=ldS, act=3, sC=(h * Lq * ldS, Lq * ldS), **qk)
    elif two_pass:
        lse = torch.empty((B, h, Lq), dtype=torch.float32, device=x.device)
        K.gemm(q.buf, k.buf, out=lse, ldc=Lk, act=5, sC=(h * Lq, Lq), **qk)
        P = torch.empty((B, h, Lq, ldS), dtype=wc.act, device=x.device)
        K.gemm(q.buf, k.buf, out=P, ldc=ldS, act=6, rowvec=lse, sC=(h * Lq * ldS, Lq * ldS), **qk)
    else:
        S = torch.empty((B, h, Lq, ldS), dtype=torch.float32, device=x.device)
        K.gemm(q.buf, k.buf, M=Lq, N=Lk, K=d, lda=q.sP, ldb=k.sP, out=S, ldc=ldS, alpha=scale, batch=(B, h),
               sA=(q.sB, q.sH), sB=(k.sB, k.sH), sC=(h * Lq * ldS, Lq * ldS), a_off=q.off, b_off=k.off)
        if spec.kind == "spatial":
            P = K.softmax_fwd(S, Lk, ldS, nq=Lq, mask_hw=thw[1] * thw[2], mask_t=thw[0], dtype=wc.act)
        else:
            P = K.softmax_fwd(S, Lk, ldS, nq=Lq, dtype=wc.act)
        del S
    Mq = B * Lq
    o = torch.empty((Mq, C), dtype=wc.act, device=x.device)
    K.gemm(P, v.buf, M=Lq, N=d, K=Lk, lda=ldS, b_kmajor=False, ldb=v.sP, out=o, ldc=C, batch=(B, h),
           sA=(h * Lq * ldS, Lq * ldS), sB=(v.sB, v.sH), sC=(Lq * C, d), b_off=v.off)
    # ---- residual path (attention.py:240 / :471) --------------------------------------------------
    arg = None
    if spec.stride_q is None or spec.kind in ("spatial", "temporal"):
        x_res = x
    elif dec:
        x_res = K.upsample_fwd(x, B, thw, C, spec.stride_q)
    else:
        assert tuple(spec.stride_q) == (1, 2, 2), "pool_skip kernel is specialised for MaxPool3d k(1,3,3) s(1,2,2)"
        x_res, arg = K.maxpool_fwd(x, B, thw, C, want_arg=save)
    rps = Lq if dp_scale is not None else 0
    x1 = K.gemm(o, wc.w(p["attn.proj.weight"]), M=Mq, N=C, K=C, bias=p["attn.proj.bias"], residual=x_res.view(Mq, C),
                out_dtype=torch.float32, row_scale=dp_a, rows_per_scale=rps)
    # ---- MLP (attention.py:243-247) -----------------------------------------------------------------
    xn2, mean2, rstd2 = K.layernorm_fwd(x1, p["norm2.weight"], p["norm2.bias"], EPS_BLOCK, out_dtype=wc.act)
    hid = spec.hidden
    # Z receives GELU'(fc1 pre-activation): the forward epilogue already evaluates it, backward only multiplies
    Z = torch.empty((Mq, hid), dtype=wc.act, device=x.device) if save else None
    hdn = K.gemm(xn2, wc.w(p["mlp.fc1.weight"]), M=Mq, N=hid, K=C, bias=p["mlp.fc1.bias"], act=1, Z=Z)
    if spec.dim != spec.dim_out:
        base = K.gemm(xn2, wc.w(p["proj.weight"]), M=Mq, N=spec.dim_out, K=C, bias=p["proj.bias"], out_dtype=torch.float32)
    else:
        base = x1
    x2 = K.gemm(hdn, wc.w(p["mlp.fc2.weight"]), M=Mq, N=spec.dim_out, K=hid, bias=p["mlp.fc2.bias"], residual=base,
                out_dtype=torch.float32, row_scale=dp_m, rows_per_scale=rps)
    y = x2.view(B, Lq, spec.dim_out)
    attn = P[..., :Lk].float() if want_attn else None
    if not save:
        return y, thw_q, None, attn
    sv.update(x=x, thw=tuple(thw), mean1=mean1, rstd1=rstd1, xn1=xn1, qkv=qkv, q=q, k=k, v=v, P=P, o=o, arg=arg,
              x1=x1, mean2=mean2, rstd2=rstd2, xn2=xn2, Z=Z, hdn=hdn, dp=dp_scale, Lq=Lq, Lk=Lk, ldS=ldS, thw_q=thw_q,
              fused_softmax=fused_softmax, two_pass=two_pass)
    return y, thw_q, sv, attn


def audio_rows(P, thw):
    """Rows of the spatial-fusion attention that SpatialAttention turns into its audio-attention map
    (av_attention.py:364-365): P[b, h, THW + t, HW*t : HW*(t+1)] for every frame t -> f32 (B, heads, T, HW)."""
    T, HW = thw[0], thw[1] * thw[2]
    return torch.stack([P[:, :, T * HW + t, HW * t: HW * (t + 1)] for t in range(T)], dim=2).float()


class _GradOut:
    """Where the parameter gradients of one block are written.

    With a GradArena (host/grad_arena.py) every gradient is a view of the block's contiguous slice of the arena,
    zeroed by ONE memset (the accumulate-into ones — norm affine, biases, pool kernels — and the split-K weight
    gradients alike).  Without one (stand-alone use of the block in tests, or a parameter that already holds a
    gradient that autograd has to accumulate into) the small gradients share one zeroed scratch buffer and the
    matrices are allocated by the GEMM."""

    def __init__(self, p, wc, dev, scratch_elems):
        arena = getattr(wc, "arena", None)
        ts = list(p.values())
        self.arena = arena if arena is not None and all(arena.has(t) and t.grad is None for t in ts) else None
        if self.arena is not None:
            self.arena.span(ts).zero_()
        else:
            self.zbuf = torch.zeros(scratch_elems, dtype=torch.float32, device=dev)
            self.zoff = 0

    def small(self, param, n):
        """zeroed f32 [n] for an accumulate-into gradient"""
        if self.arena is not None:
            return self.arena.view(param).view(-1)
        lo = self.zoff
        self.zoff = lo + (n + 3) // 4 * 4          # keep 16-byte alignment for the vectorised kernels
        return self.zbuf[lo: lo + n]

    def matrix(self, param):
        """(zeroed output tensor, or None when the GEMM should allocate its own)"""
        return self.arena.view(param) if self.arena is not None else None


class _Fork:
    """Weight-gradient work of a block runs on a second stream: it is off the dX critical path, and two thirds of
    the step's GEMMs fill at most one wave of SMs, so the two streams share the machine.  Inside the captured
    step the fork becomes a parallel branch of the CUDA graph.  Every tensor the side stream reads is kept alive
    in `hold` until join() — the caching allocator is not told about the second stream."""

    def __init__(self, stream=None):
        self.side = stream
        self.hold = []
        self.used = False

    def run(self, fn, *keep):
        if self.side is None:
            return fn()
        self.hold.extend(keep)
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            r = fn()
        self.used = True
        return r

    def join(self, waiter=None):
        """Make `waiter` (default: the current stream) wait for the branch; drop the kept tensors."""
        if self.used:
            (waiter if waiter is not None else torch.cuda.current_stream()).wait_stream(self.side)
        if waiter is None:
            self.used = False
            self.hold.clear()


def block_backward(spec, p, wc, sv, dy, d_audio_rows=None, hand=None):
    """dy: f32 (B, Lq, dim_out).  Returns (dx f32 (B,N,C), {param name: grad}).  d_audio_rows: gradient w.r.t.
    audio_rows(P) (MVIT.SPATIAL_AUDIO_ATTN), folded into dP before the softmax backward.
    hand = {"dp": MLP-branch DropPath scale of the block that produced x, or None}: the closing LayerNorm backward then also
    writes the 16-bit, DropPath-scaled copy of dx that the producer's backward starts from (its `g2`), handed over
    through wc.handoff — one cast launch per block less."""
    x = sv["x"]
    B, N, C = x.shape
    h, d = spec.heads, spec.head_dim
    dec = spec.kind == "dec"
    Lq, Lk, ldS, thw, thw_q = sv["Lq"], sv["Lk"], sv["ldS"], sv["thw"], sv["thw_q"]
    M, Mq, hid, Co = B * N, B * Lq, spec.hidden, spec.dim_out
    dp = sv["dp"]
    dp_a, dp_m = (dp[0], dp[1]) if dp is not None else (None, None)
    rps = Lq if dp is not None else 0
    dev = x.device
    g = {}

    n_pool = 3 if spec.stride_q is not None and spec.stride_kv is not None else (2 if spec.stride_kv is not None else
                                                                               (1 if spec.stride_q is not None else 0))
    go = _GradOut(p, wc, dev, 4 * C + 3 * C + C + hid + 2 * Co + n_pool * 29 * d + 64)
    # the second stream needs gradient storage that outlives the block (the arena); see _Fork
    fork = wc.backward_fork() if go.arena is not None and hasattr(wc, "backward_fork") else _Fork()

    def zeros(name, n):
        return go.small(p[name], n)

    def wgrad(a_rows, b_rows, wname, m_out, n_out, ktok, bias_name):
        # dW[m_out, n_out] = a_rows^T . b_rows   (both token-major: contraction over rows); the bias gradient
        # sum_rows a_rows comes out of the same kernel (one extra narrow MMA per k-step against an all-ones tile)
        g[bias_name] = zeros(bias_name, m_out)
        out = go.matrix(p[wname])
        g[wname] = fork.run(lambda: K.gemm(a_rows, b_rows, M=m_out, N=n_out, K=ktok, a_kmajor=False, b_kmajor=False, lda=m_out,
                                           ldb=n_out, out=None if out is None else out.view(m_out, n_out), out_dtype=torch.float32,
                                           out_is_zero=out is not None, split_k=-1, rowsum=g[bias_name]),
                            a_rows, b_rows)

    dy = dy.contiguous().view(Mq, Co)
    g2 = wc.take_handoff(dy) if hasattr(wc, "take_handoff") else None         # written by the consumer block's LayerNorm backward
    if g2 is None:
        g2 = K.cast16(dy, wc.grad, row_scale=dp_m, rows_per_scale=rps)        # gradient entering the (drop-path scaled) MLP branch
    # ---- fc2, GELU, fc1 ----------------------------------------------------------------------------
    wgrad(g2, sv["hdn"], "mlp.fc2.weight", Co, hid, Mq, "mlp.fc2.bias")
    dZ = K.gemm(g2, wc.w(p["mlp.fc2.weight"]), M=Mq, N=hid, K=Co, b_kmajor=False, act=2, Z=sv["Z"])
    wgrad(dZ, sv["xn2"], "mlp.fc1.weight", hid, C, Mq, "mlp.fc1.bias")
    dxn2 = K.gemm(dZ, wc.w(p["mlp.fc1.weight"]), M=Mq, N=C, K=hid, b_kmajor=False)
    del dZ
    g["norm2.weight"], g["norm2.bias"] = zeros("norm2.weight", C), zeros("norm2.bias", C)
    if spec.dim != spec.dim_out:
        gp = g2 if dp is None else K.cast16(dy, wc.grad)                     # the re-based residual is not drop-path scaled
        wgrad(gp, sv["xn2"], "proj.weight", Co, C, Mq, "proj.bias")
        K.gemm(gp, wc.w(p["proj.weight"]), M=Mq, N=C, K=Co, b_kmajor=False, out=dxn2, accumulate=True)
        dx1, g1 = K.layernorm_bwd(dxn2, sv["x1"], sv["mean2"], sv["rstd2"], p["norm2.weight"], g["norm2.weight"], g["norm2.bias"],
                                  copy16=wc.grad, row_scale=dp_a, rows_per_scale=rps)
    else:
        dx1, g1 = K.layernorm_bwd(dxn2, sv["x1"], sv["mean2"], sv["rstd2"], p["norm2.weight"], g["norm2.weight"], g["norm2.bias"],
                                  add=dy, copy16=wc.grad, row_scale=dp_a, rows_per_scale=rps)
    del dxn2, g2
    # ---- attention output projection (g1 = drop-path scaled 16-bit copy of dx1, written by the LayerNorm backward) ------
    wgrad(g1, sv["o"], "attn.proj.weight", C, C, Mq, "attn.proj.bias")
    do = K.gemm(g1, wc.w(p["attn.proj.weight"]), M=Mq, N=C, K=C, b_kmajor=False)       # (B, Lq, heads, d)
    del g1
    # ---- residual path -----------------------------------------------------------------------------------
    if spec.stride_q is None or spec.kind in ("spatial", "temporal"):
        dx_skip = dx1.view(B, N, C)
    elif dec:
        dx_skip = K.upsample_bwd(dx1, B, thw, C, spec.stride_q)
    else:
        dx_skip = K.maxpool_bwd(dx1, sv["arg"], B, thw, C)
    # ---- attention backward ------------------------------------------------------------------------------
    q, k, v, P = sv["q"], sv["k"], sv["v"], sv["P"]
    dqkv = torch.empty((M, 3 * C), dtype=wc.grad, device=dev)
    qs = (N * 3 * C, d, 3 * C)
    pooled_q, pooled_kv = spec.stride_q is not None, spec.stride_kv is not None

    # dV = P^T.dO and dK = dS^T.Q contract over the Lq queries.  With few pooled keys and many queries (the decoder: 64 keys,
    # up to 32768 queries) the product has fewer output tiles than SMs: it is then split over the queries into an f32
    # buffer (vectorised reduce-adds), which the LayerNorm backward of the pool reads directly.
    split_kv = pooled_kv and B * h * ((Lk + 127) // 128) * 2 <= 148 and Lq >= 2048

    def grad_target(pooled, L, slot, split=False):
        if pooled and split:
            t = torch.zeros((B, h, L, d), dtype=torch.float32, device=dev)
            return t, dict(out=t, ldc=d, sC=(h * L * d, L * d), c_off=0, out_is_zero=True, split_k=-1)
        if pooled:
            t = torch.empty((B, h, L, d), dtype=wc.grad, device=dev)
            return t, dict(out=t, ldc=d, sC=(h * L * d, L * d), c_off=0)
        return None, dict(out=dqkv, ldc=3 * C, sC=(qs[0], qs[1]), c_off=slot * C)

    sP = (h * Lq * ldS, Lq * ldS)
    dv_t, tgt = grad_target(pooled_kv, Lk, 2, split_kv)
    K.gemm(P, do, M=Lk, N=d, K=Lq, a_kmajor=False, lda=ldS, b_kmajor=False, ldb=C, batch=(B, h), sA=sP, sB=(Lq * C, d), **tgt)
    if sv["fused_softmax"]:
        # dS = scale * P o (dP - rowsum(dP o P)) in the epilogue of the dO.v^T kernel: dP never reaches HBM
        dS = torch.empty(P.shape, dtype=wc.grad, device=dev)
        K.gemm(do, v.buf, M=Lq, N=Lk, K=d, lda=C, ldb=v.sP, out=dS, ldc=ldS, alpha=d ** -0.5, act=4, Z=P, batch=(B, h),
               sA=(Lq * C, d), sB=(v.sB, v.sH), sC=sP, b_off=v.off)
    elif sv.get("two_pass") and d_audio_rows is None:
        # rowsum(dP o P) = dO . O: with that row term in hand dS = scale * P o (dP - D) is element-wise in the epilogue of
        # the dO.v^T kernel, for any number of keys — dP never reaches HBM
        D = K.rowdot(do, sv["o"], B, Lq, h, d)
        dS = torch.empty(P.shape, dtype=wc.grad, device=dev)
        K.gemm(do, v.buf, M=Lq, N=Lk, K=d, lda=C, ldb=v.sP, out=dS, ldc=ldS, alpha=d ** -0.5, act=7, Z=P, rowvec=D, batch=(B, h),
               sA=(Lq * C, d), sB=(v.sB, v.sH), sC=sP, b_off=v.off)
    else:
        dP = torch.empty((B, h, Lq, ldS), dtype=torch.float32, device=dev)
        K.gemm(do, v.buf, M=Lq, N=Lk, K=d, lda=C, ldb=v.sP, out=dP, ldc=ldS, batch=(B, h), sA=(Lq * C, d), sB=(v.sB, v.sH), sC=sP,
               b_off=v.off)
        if d_audio_rows is not None:
            T_, HW_ = thw[0], thw[1] * thw[2]
            for t in range(T_):
                dP[:, :, T_ * HW_ + t, HW_ * t: HW_ * (t + 1)] += d_audio_rows[:, :, t]
        dS = K.softmax_bwd(P, dP, Lk, d ** -0.5, dtype=wc.grad)
        del dP
    dq_t, tgt = grad_target(pooled_q, Lq, 0)
    K.gemm(dS, k.buf, M=Lq, N=d, K=Lk, lda=ldS, b_kmajor=False, ldb=k.sP, batch=(B, h), sA=sP, sB=(k.sB, k.sH), b_off=k.off, **tgt)
    dk_t, tgt = grad_target(pooled_kv, Lk, 1, split_kv)
    K.gemm(dS, q.buf, M=Lk, N=d, K=Lq, a_kmajor=False, lda=ldS, b_kmajor=False, ldb=q.sP, batch=(B, h), sA=sP, sB=(q.sB, q.sH),
           b_off=q.off, **tgt)
    del dS

    def pool_ln_backward(dt, key, nname):
        pre, mean, rstd = sv[key]
        g[nname + ".weight"], g[nname + ".bias"] = zeros(nname + ".weight", d), zeros(nname + ".bias", d)
        return K.layernorm_bwd(dt, pre, mean, rstd, p[nname + ".weight"], g[nname + ".weight"], g[nname + ".bias"], dx_dtype=wc.grad)

    def pool_backward(items, stride, transposed, grid_out):
        """items: [(du, slot, weight name)] — one pool, or the k and v pools of the block in a single launch each for the
        data gradient (adjoint gather written straight into the qkv-gradient slice) and the weight gradient."""
        L = grid_out[0] * grid_out[1] * grid_out[2]
        dense = (h * L * d, L * d, d)
        (du, slot, wname) = items[0]
        du2 = slot2 = wname2 = None
        for _, _, wn in items:
            g[wn] = zeros(wn, 27 * d).view_as(p[wn])
        if len(items) == 2:
            du2, slot2, wname2 = items[1]

        def weight_grads():
            if transposed:
                assert len(items) == 1
                K.dwconv_wgrad(sv["qkv"], qs, slot * C, thw, du, dense, 0, grid_out, B, h, d, stride, g[wname])
            else:
                sec = None
                if len(items) == 2:
                    sec = dict(small=du2, small_off=0, big=sv["qkv"], big_off=slot2 * C, dw=g[wname2])
                K.dwconv_wgrad(du, dense, 0, grid_out, sv["qkv"], qs, slot * C, thw, B, h, d, stride, g[wname], second=sec)

        fork.run(weight_grads, du, du2, sv["qkv"])
        sec = None
        if len(items) == 2:
            sec = dict(inp=du2, in_off=0, w=p[wname2], out=dqkv, out_off=slot2 * C)
        K.dwconv(du, dense, 0, B, h, d, grid_out, stride, p[wname], transposed=not transposed, out=dqkv, out_strides=qs,
                 out_off=slot * C, thw_out=thw, second=sec)

    if pooled_q:
        qname = "attn.upsample_q.weight" if dec else "attn.pool_q.weight"
        pool_backward([(pool_ln_backward(dq_t, "q_pool", "attn.norm_q"), 0, qname)], spec.stride_q, dec, thw_q)
    if pooled_kv:
        # norm_k and norm_v backward share their geometry: one launch
        for nn_ in ("attn.norm_k", "attn.norm_v"):
            g[nn_ + ".weight"], g[nn_ + ".bias"] = zeros(nn_ + ".weight", d), zeros(nn_ + ".bias", d)
        (kp, km, kr), (vp, vm, vr) = sv["k_pool"], sv["v_pool"]
        duk, duv = K.layernorm_bwd_pair((dk_t, dv_t), (kp, vp), (km, vm), (kr, vr), (p["attn.norm_k.weight"], p["attn.norm_v.weight"]),
                                        (g["attn.norm_k.weight"], g["attn.norm_v.weight"]), (g["attn.norm_k.bias"], g["attn.norm_v.bias"]),
                                        dx_dtype=wc.grad)
        pool_backward([(duk, 1, "attn.pool_k.weight"), (duv, 2, "attn.pool_v.weight")], spec.stride_kv, False, k.thw)
    # ---- qkv projection and norm1 ---------------------------------------------------------------------------
    wgrad(dqkv, sv["xn1"].view(M, C), "attn.qkv.weight", 3 * C, C, M, "attn.qkv.bias")
    dxn1 = K.gemm(dqkv, wc.w(p["attn.qkv.weight"]), M=M, N=C, K=3 * C, b_kmajor=False)
    g["norm1.weight"], g["norm1.bias"] = zeros("norm1.weight", C), zeros("norm1.bias", C)
    if hand is not None and hasattr(wc, "give_handoff"):
        hdp = hand.get("dp")
        dx, dx16 = K.layernorm_bwd(dxn1, x, sv["mean1"], sv["rstd1"], p["norm1.weight"], g["norm1.weight"], g["norm1.bias"],
                                   add=dx_skip.view(B, N, C), copy16=wc.grad, row_scale=None if hdp is None else hdp[1],
                                   rows_per_scale=N if hdp is not None else 0)
        wc.give_handoff(dx, dx16.view(M, C))
    else:
        dx = K.layernorm_bwd(dxn1, x, sv["mean1"], sv["rstd1"], p["norm1.weight"], g["norm1.weight"], g["norm1.bias"],
                             add=dx_skip.view(B, N, C))
    if not getattr(wc, "defer_join", False):
        fork.join()         # by default every block hands complete gradients to autograd (DDP's reducer, plain optimizers)
    return dx, g


class BlockFn(torch.autograd.Function):
    """autograd node: (x, *block parameters) -> block output.  Non-tensor context (spec, weight cache,
    grid) travels in `meta`."""

    @staticmethod
    def forward(ctx, meta, x, *params):
        spec, wc, thw, dp_scale, names = meta[:5]
        extra = meta[5] if len(meta) > 5 else None       # {"want": "attn" | "audio_rows"}: optional attention outputs
        ctx.hand = meta[6] if len(meta) > 6 else None    # gradient hand-over to the producer block (block_backward)
        p = dict(zip(names, params))
        need = any(ctx.needs_input_grad)      # forward runs under no_grad: ask the node, not the mode
        # MODEL.ACT_CHECKPOINT (custom_multimodal_builder.py:154-155,178-179,214-215: fairscale checkpoint_wrapper around the
        # video and audio encoder blocks): keep only the block input, re-run the forward inside backward
        ckpt = need and getattr(wc, "act_checkpoint", False) and spec.kind == "enc" and extra is None
        x = x.contiguous()
        y, thw_q, sv, attn = block_forward(spec, p, wc, x, thw, dp_scale, save=need and not ckpt, want_attn=extra is not None)
        ctx.meta = meta
        ctx.sv = sv
        ctx.recompute = (x, tuple(thw), dp_scale) if ckpt else None
        ctx.params = params
        ctx.thw_q = thw_q
        if extra is not None and extra.get("want") == "audio_rows":
            assert spec.kind == "spatial"
            return y, audio_rows(attn, thw)              # second, differentiable output
        if extra is not None:
            extra["attn"] = attn                         # visualisation output (detached)
        return y

    @staticmethod
    def backward(ctx, dy, d_audio_rows=None):
        spec, wc, thw, dp_scale, names = ctx.meta[:5]
        p = dict(zip(names, ctx.params))
        if dy is None:
            dy = torch.zeros(ctx.sv["x"].shape[0], ctx.sv["Lq"], spec.dim_out, dtype=torch.float32, device=ctx.sv["x"].device)
        if ctx.recompute is not None:
            x, thw_in, dp = ctx.recompute
            _, _, ctx.sv, _ = block_forward(spec, p, wc, x, thw_in, dp, save=True)
            ctx.recompute = None
        dx, g = block_backward(spec, p, wc, ctx.sv, dy, d_audio_rows, hand=ctx.hand)
        ctx.sv = None
        return (None, dx) + tuple(g[n].view_as(p[n]) for n in names)
