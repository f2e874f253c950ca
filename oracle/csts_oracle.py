"""CPU oracle for the CSTS hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional fp32 restatement (plain ``torch.nn.functional`` on a flat ``{name: tensor}`` state
dict) of the reference's forward pass and ``kldiv+egonce`` loss.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import this file; nothing under ``csts_b200/`` does, and the product path never falls back to
it.

Pinning: the reference ships NO tests, golden vectors or fixtures for this path
 so parity is unpinned *by the reference's own tests*.  This oracle is
instead pinned against the reference itself: ``oracle/make_golden.py`` imports the unmodified
reference (through ``oracle/ref_shim.py``) in the build container, runs it on seeded inputs and
commits the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file
against those fixtures (and against the live reference whenever /root/reference is present).

All ``ref:`` citations are paths relative to the reference root (BolinLai/CSTS).
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# Architecture table.  ref: slowfast/models/custom_multimodal_builder.py:107-300 evaluated with
# configs/{Ego4D,Aria}/*.yaml (all four YAMLs give the same numbers).
# Each row: (prefix, kind, dim, dim_out, heads, stride_q, stride_kv)
#   kind: "enc" MultiScaleBlock | "dec" MultiScaleDecoderBlock | "spatial" | "temporal"
# --------------------------------------------------------------------------------------------


def _round_width(width, mult, divisor):
    # ref: slowfast/models/utils.py:8-21 (round_width with min_width=1)
    width = width * mult
    out = max(1, int(width + divisor / 2) // divisor * divisor)
    if out < 0.9 * width:
        out += divisor
    return int(out)


def arch_table(depth=16, embed_dim=96, num_heads=1,
               dim_mul=((1, 2.0), (3, 2.0), (14, 2.0)), head_mul=((1, 2.0), (3, 2.0), (14, 2.0)),
               pool_q_stride=((1, 1, 2, 2), (3, 1, 2, 2), (14, 1, 2, 2)), kv_adaptive=(1, 8, 8)):
    dm = [1.0] * (depth + 1)
    hm = [1.0] * (depth + 1)
    for i, m in dim_mul:
        dm[i] = m
    for i, m in head_mul:
        hm[i] = m
    sq = {r[0]: tuple(r[1:]) for r in pool_q_stride}
    rows = []
    skv = list(kv_adaptive)
    heads, dim = num_heads, embed_dim
    for i in range(depth):
        # ref: custom_multimodal_builder.py:129-136 (adaptive kv stride) and :150-153 (widths)
        if i in sq:
            skv = [max(skv[d] // sq[i][d], 1) for d in range(3)]
        heads = _round_width(heads, hm[i], 1)
        dim = _round_width(dim, dm[i], heads)
        dim_out = _round_width(dim, dm[i + 1], _round_width(heads, hm[i + 1], 1))
        rows.append((f"blocks.{i}", "enc", dim, dim_out, heads, sq.get(i), tuple(skv)))
    # audio encoder, ref: custom_multimodal_builder.py:184-191
    a_dim, a_out, a_heads = [96, 192, 384, 768], [192, 384, 768, 768], [1, 2, 4, 8]
    a_sq = [None, (1, 2, 2), (1, 2, 2), (1, 2, 2)]
    a_skv = [(1, 8, 8), (1, 4, 4), (1, 2, 2), (1, 1, 1)]
    for i in range(4):
        rows.append((f"blocks_audio.{i}", "enc", a_dim[i], a_out[i], a_heads[i], a_sq[i], a_skv[i]))
    tok = rows[depth - 1][3]
    rows.append(("spatial_fusion", "spatial", tok, tok, heads, None, None))     # ref :253-270
    rows.append(("temporal_fusion", "temporal", tok, tok, heads, None, None))   # ref :232-249
    # decoder, ref: custom_multimodal_builder.py:272-281
    d_in, d_out, d_heads = [768, 768, 384, 192], [768, 384, 192, 96], [8, 4, 4, 2]
    d_sq = [(1, 2, 2), (1, 2, 2), (1, 2, 2), (2, 1, 1)]
    d_skv = [(1, 2, 2), (1, 4, 4), (1, 8, 8), (1, 16, 16)]
    for i in range(4):
        rows.append((f"decode_block{i + 1}", "dec", d_in[i], d_out[i], d_heads[i], d_sq[i], d_skv[i]))
    return {r[0]: r for r in rows}


ARCH = arch_table()

# --------------------------------------------------------------------------------------------
# Optional bf16 emulation.  With EMULATE_BF16 = True the oracle inserts round-to-bf16 at exactly the
# tensors that csts_b200 stores in bf16 (GEMM operands, saved activations and the operand
# gradients of backward), while accumulation, the residual stream, statistics and parameter
# gradients stay f32 — the precision contract of DESIGN.md §3.  Comparing the CUDA path with this
# variant separates *implementation* error (must be ~1e-3) from *rounding-policy* error (what the
# plain fp32 oracle measures).  Off by default: the oracle proper is fp32.
# --------------------------------------------------------------------------------------------
EMULATE_BF16 = False
FWD_DTYPE = torch.bfloat16     # storage type of the forward activations / weight copies (fp16 in the MIXED_PRECISION mode)
BWD_DTYPE = torch.bfloat16     # storage type of the activation gradients


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _fw(t):
    return t.to(FWD_DTYPE).to(torch.float32)


def _bw(t):
    return t.to(BWD_DTYPE).to(torch.float32)


class _RoundBoth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return _fw(t)

    @staticmethod
    def backward(ctx, g):
        return _bw(g)


class _RoundBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return _bw(g)


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return _fw(t)

    @staticmethod
    def backward(ctx, g):
        return g


class _GeluSavedBf16(torch.autograd.Function):
    """gelu(z) in f32 whose backward evaluates gelu' at the bf16-rounded pre-activation (the saved Z)."""

    @staticmethod
    def forward(ctx, z):
        ctx.save_for_backward(_fw(z))
        return F.gelu(z)

    @staticmethod
    def backward(ctx, g):
        (zb,) = ctx.saved_tensors
        cdf = 0.5 * (1 + torch.erf(zb * 0.7071067811865476))
        pdf = 0.3989422804014327 * torch.exp(-0.5 * zb * zb)
        return g * (cdf + zb * pdf)


class _SoftmaxSavedBf16(torch.autograd.Function):
    """softmax whose output is stored in bf16 and whose backward uses that stored value."""

    @staticmethod
    def forward(ctx, s):
        p = _fw(s.softmax(dim=-1))
        ctx.save_for_backward(p)
        return p

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        return p * (g - (g * p).sum(-1, keepdim=True))


SKIP = set()     # precision study (oracle/precision_study.py): "f:<tag>" / "b:<tag>" disables one rounding point


def rnd(t, tag="x"):      # stored in bf16, gradient stored in bf16
    if not EMULATE_BF16:
        return t
    f, b = ("f:" + tag) not in SKIP, ("b:" + tag) not in SKIP
    if f and b:
        return _RoundBoth.apply(t)
    if f:
        return _RoundFwd.apply(t)
    if b:
        return _RoundBwd.apply(t)
    return t


def rnd_f(t, tag="x"):    # stored in bf16 (forward only)
    return _RoundFwd.apply(t) if EMULATE_BF16 and ("f:" + tag) not in SKIP else t


def rnd_b(t, tag="x"):    # gradient arriving here is cast to bf16 before it feeds the backward GEMMs
    return _RoundBwd.apply(t) if EMULATE_BF16 and ("b:" + tag) not in SKIP else t


def _gelu(z):
    return _GeluSavedBf16.apply(z) if EMULATE_BF16 else F.gelu(z)


def _softmax(s):
    return _SoftmaxSavedBf16.apply(s) if EMULATE_BF16 else s.softmax(dim=-1)
EPS_BLOCK = 1e-6   # norm1/norm2: partial(nn.LayerNorm, eps=1e-6), ref custom_multimodal_builder.py:61
EPS_POOL = 1e-5    # norm_q/k/v: plain nn.LayerNorm, ref attention.py:206


# --------------------------------------------------------------------------------------------
# Token-grid helpers
# --------------------------------------------------------------------------------------------

def _to_grid(t, thw):
    """(B,h,L,d) tokens -> (B*h, d, T, H, W).  ref: attention.py:29-31"""
    B, h, L, d = t.shape
    T, H, W = thw
    return t.reshape(B * h, T, H, W, d).permute(0, 4, 1, 2, 3)


def _from_grid(g, B, h):
    """(B*h, d, T', H', W') -> (B,h,L',d), thw'.  ref: attention.py:35-37"""
    d = g.shape[1]
    thw = (g.shape[2], g.shape[3], g.shape[4])
    return g.reshape(B, h, d, -1).transpose(2, 3), thw


def pool_tokens(t, thw, w, stride, ln_w, ln_b, transposed=False):
    """Depthwise 3x3x3 conv (or transposed conv) over the token grid, then LayerNorm(head_dim).

    ref: attention_pool attention.py:11-49 with pool = Conv3d(d,d,3,stride,pad=1,groups=d,
    bias=False) (:105-116); attention_upsample :251-292 with ConvTranspose3d(...,
    output_padding=stride-1) (:344-348).  LayerNorm eps is the torch default 1e-5.
    """
    B, h, L, d = t.shape
    g = _to_grid(t, thw)
    w = rnd_f(w, "poolw")       # storage emulation: the kernel in the activations' 16-bit type (conv3d under autocast; the CUDA
                                # path's mixed-precision FMAs take both factors in 16 bits); its gradient stays f32
    if transposed:
        op = tuple(0 if s == 1 else s - 1 for s in stride)
        g = F.conv_transpose3d(g, w, None, stride=stride, padding=1, output_padding=op, groups=d)
    else:
        g = F.conv3d(g, w, None, stride=stride, padding=1, groups=d)
    t, thw = _from_grid(g, B, h)
    t = rnd(F.layer_norm(rnd(t,"pre"), (d,), ln_w, ln_b, EPS_POOL),"pooled")
    return t, thw


def spatial_mask(thw, device):
    """In-frame additive offset for SpatialAttention.  ref: av_attention.py:336-345.

    Token (t, .) may attend to frame t's H*W visual tokens and to audio token t; every other
    logit gets -1e8.
    """
    T, HW = thw[0], thw[1] * thw[2]
    n = T * HW + T
    frame = torch.cat([torch.arange(T).repeat_interleave(HW), torch.arange(T)]).to(device)
    allowed = frame[:, None] == frame[None, :]
    off = torch.full((n, n), 1e8, device=device)
    off[allowed] = 0.0
    return off


def attention(sd, pfx, x, thw, heads, stride_q, stride_kv, kind, want_attn=False):
    """MultiScaleAttention / MultiScaleDecoderAttention / Spatial / TemporalAttention forward.

    ref: attention.py:120-162, :365-392; av_attention.py:120-152, :322-372.
    """
    B, N, C = x.shape
    d = C // heads
    qkv = rnd(F.linear(x, rnd_f(sd[pfx + "qkv.weight"],"w"), sd[pfx + "qkv.bias"]),"qkv")
    qkv = qkv.reshape(B, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q_thw = thw
    if kind == "dec":
        q, q_thw = pool_tokens(q, thw, sd[pfx + "upsample_q.weight"], stride_q,
                               sd[pfx + "norm_q.weight"], sd[pfx + "norm_q.bias"], transposed=True)
    elif stride_q is not None:
        q, q_thw = pool_tokens(q, thw, sd[pfx + "pool_q.weight"], stride_q,
                               sd[pfx + "norm_q.weight"], sd[pfx + "norm_q.bias"])
    if stride_kv is not None:
        k, _ = pool_tokens(k, thw, sd[pfx + "pool_k.weight"], stride_kv,
                           sd[pfx + "norm_k.weight"], sd[pfx + "norm_k.bias"])
        v, _ = pool_tokens(v, thw, sd[pfx + "pool_v.weight"], stride_kv,
                           sd[pfx + "norm_v.weight"], sd[pfx + "norm_v.bias"])
    s = rnd_b(q @ k.transpose(-2, -1),"dS") * (d ** -0.5)
    if kind == "spatial":
        s = s - spatial_mask(thw, s.device)
    p = _softmax(s)
    o = rnd((p @ v).transpose(1, 2).reshape(B, q.shape[2], C),"o")
    o = rnd_b(F.linear(o, rnd_f(sd[pfx + "proj.weight"],"w"), sd[pfx + "proj.bias"]),"g1")
    return (o, q_thw, p) if want_attn else (o, q_thw)


def skip_path(x, thw, kind, stride_q):
    """Residual-path resampling.  ref: attention.py:225-236,240 (MaxPool3d k=s+1, p=k//2) and
    :463-467,471 (nn.Upsample trilinear, align_corners=False)."""
    if stride_q is None or kind in ("spatial", "temporal"):
        return x
    B, N, C = x.shape
    g = _to_grid(x.unsqueeze(1), thw)
    if kind == "dec":
        g = F.interpolate(g, scale_factor=tuple(float(s) for s in stride_q), mode="trilinear")
    else:
        ks = tuple(s + 1 if s > 1 else s for s in stride_q)
        g = F.max_pool3d(g, ks, stride_q, tuple(k // 2 for k in ks))
    t, _ = _from_grid(g, B, 1)
    return t.squeeze(1)


def block(sd, name, x, thw, want_attn=False, spec=None, drop=None):
    """One transformer block (any of the four kinds).  ref: attention.py:238-248, :469-479;
    av_attention.py:229-250, :450-473.  DropPath (ref common.py:46-59) is identity unless `drop` = (attention-branch
    scale (B,), MLP-branch scale (B,)) is given: the two per-sample factors mask / keep_prob that the reference draws
    with torch.rand at attention.py:242 and :247, supplied by the caller so a run can be reproduced exactly.
    `spec` = (kind, dim, dim_out, heads, stride_q, stride_kv) overrides the ARCH row (unit tests)."""
    kind, dim, dim_out, heads, sq, skv = spec if spec is not None else ARCH[name][1:]
    p = name + "."
    da = dm = None
    if drop is not None:
        da, dm = (t.reshape(-1, 1, 1).to(x.dtype) for t in drop)
    xn = rnd(F.layer_norm(x, (dim,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], EPS_BLOCK),"xn")
    res = attention(sd, p + "attn.", xn, thw, heads, sq, skv, kind, want_attn)
    x = skip_path(x, thw, kind, sq) + (res[0] if da is None else res[0] * da)
    xn = rnd(F.layer_norm(x, (dim,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], EPS_BLOCK),"xn")
    # decoder MLP hidden is 4*dim_out (ref attention.py:444) — implied by the weight shapes
    hid = rnd_f(_gelu(rnd_b(F.linear(xn, rnd_f(sd[p + "mlp.fc1.weight"],"w"), sd[p + "mlp.fc1.bias"]),"dZ")),"h")     # exact erf
    mlp = rnd_b(F.linear(hid, rnd_f(sd[p + "mlp.fc2.weight"],"w"), sd[p + "mlp.fc2.bias"]),"g2")
    if dim != dim_out:
        x = rnd_b(F.linear(xn, rnd_f(sd[p + "proj.weight"],"wproj"), sd[p + "proj.bias"]),"gproj")    # ref :245-246
    x = x + (mlp if dm is None else mlp * dm)
    return (x, res[1], res[2]) if want_attn else (x, res[1])


def patch_embed(sd, name, x):
    """ref: stem_helper.py:35-38 — Conv3d k(3,7,7) s(2,4,4) p(1,3,3), flatten, transpose."""
    y = rnd_b(F.conv3d(rnd_f(x,"patchx"), rnd_f(sd[name + ".proj.weight"],"patchw"), sd[name + ".proj.bias"], stride=(2, 4, 4), padding=(1, 3, 3)),"patchg")
    return y.flatten(2).transpose(1, 2)


def sep_pos_embed(spatial, temporal):
    """ref: custom_multimodal_builder.py:362-365 — spatial.repeat(T) + repeat_interleave(temporal)."""
    T, HW = temporal.shape[1], spatial.shape[1]
    return spatial.repeat(1, T, 1) + temporal.repeat_interleave(HW, dim=1)


def frame_pool(sd, name, tok, thw):
    """Dense Conv3d(768,768,(1,8,8)) over a (B, T*8*8, C) token map -> (B, T, C).
    ref: custom_multimodal_builder.py:227-229, :420-421."""
    B, N, C = tok.shape
    g = rnd_f(tok).reshape(B, *thw, C).permute(0, 4, 1, 2, 3)
    y = rnd_b(F.conv3d(g, rnd_f(sd[name + ".weight"]), sd[name + ".bias"]))
    return y.squeeze(-1).squeeze(-1).permute(0, 2, 1)


def audio_rescale(p, thw):
    """SpatialAttention's audio-attention map (ref av_attention.py:360-370): the attention of audio token t over
    the H*W visual tokens of frame t, min-max rescaled per (b, head, t).  p (B,h,THW+T,THW+T) -> (B,h,T,H,W)."""
    T, H, W = thw
    HW, THW = H * W, T * H * W
    a = torch.stack([p[:, :, THW + t, HW * t: HW * (t + 1)] for t in range(T)], dim=2)
    amax = a.max(dim=-1, keepdim=True)[0]
    amin = a.min(dim=-1, keepdim=True)[0]
    a = (a - amin) / (amax - amin + 1e-8)
    return a.reshape(a.shape[0], a.shape[1], T, H, W)


def csts_forward(sd, video, audio, return_embed=False, return_intermediates=False, spatial_audio_attn=False,
                 return_spatial_attn=False, return_temporal_attn=False, drop_scales=None):
    """CSTS.forward.  ref: custom_multimodal_builder.py:343-498 (CLS_EMBED_ON False, SEP_POS_EMBED True,
    dropout 0; `spatial_audio_attn` = MVIT.SPATIAL_AUDIO_ATTN, default False in every shipped YAML).

    video (B,3,8,256,256), audio (B,1,8,256,256) -> logits (B,1,8,64,64) [, v (B,256), a (B,256)]
    With return_spatial_attn / return_temporal_attn (and no return_embed): [logits, spatial_attn?, temporal_attn?]
    (ref :483-491), the attention probabilities of the two fusion blocks.
    drop_scales: {block name: (attention scale (B,), MLP scale (B,))} — DropPath factors of a training forward.
    """
    inter = {}
    drop_scales = drop_scales or {}
    x = patch_embed(sd, "patch_embed", video)
    y = patch_embed(sd, "patch_embed_audio", audio)
    B = x.shape[0]
    T = video.shape[2] // 2
    H = video.shape[3] // 4
    W = video.shape[4] // 4
    x = x + sep_pos_embed(sd["pos_embed_spatial"], sd["pos_embed_temporal"])
    y = y + sep_pos_embed(sd["pos_embed_spatial_audio"], sd["pos_embed_temporal_audio"])
    thw, thw_a = (T, H, W), (T, H, W)
    skips = [(x, thw)]
    inter["stem_video"], inter["stem_audio"] = x, y
    for i in range(16):                                   # ref :386-411 (interleaving is cosmetic)
        x, thw = block(sd, f"blocks.{i}", x, thw, drop=drop_scales.get(f"blocks.{i}"))
        if i in (0, 2, 13):
            skips.append((x, thw))
    for i in range(4):
        y, thw_a = block(sd, f"blocks_audio.{i}", y, thw_a)
    inter["enc_video"], inter["enc_audio"] = x, y
    # spatial fusion, ref :414-432
    y_sp = frame_pool(sd, "audio_pool", y, thw_a)
    av, _, p_sp = block(sd, "spatial_fusion", torch.cat([x, y_sp], dim=1), thw, want_attn=True)
    x_sp = av[:, : x.shape[1]]
    # temporal fusion, ref :435-451
    x_in = x
    if spatial_audio_attn:                                  # ref :438-440: x_temporal * mean_heads(audio_rescale)
        w_a = audio_rescale(p_sp, thw).mean(dim=1).reshape(B, -1, 1)
        x_in = x * w_a
    x_t = frame_pool(sd, "vision_pool", x_in, thw)
    y_t = frame_pool(sd, "audio_pool2", y, thw_a)
    av_t, _, p_tm = block(sd, "temporal_fusion", torch.cat([x_t, y_t], dim=1), (2, 2, 2), want_attn=True)
    # re-weight, ref :454-461
    C = x.shape[2]
    nt = x_t.shape[1]
    xw = (x_sp.reshape(B, *thw, C) * av_t[:, :nt, None, None, :]).reshape(B, -1, C)
    yw = (y.reshape(B, *thw_a, C) * av_t[:, nt:, None, None, :]).reshape(B, -1, C)
    inter["x_reweight"], inter["y_reweight"] = xw, yw
    # decoder, ref :465-479
    f = xw
    for i in range(4):
        f, thw = block(sd, f"decode_block{i + 1}", f, thw)
        if i < 3:
            f = f + skips[3 - i][0]
    inter["decoded"] = f
    f = f.reshape(B, *thw, f.shape[2]).permute(0, 4, 1, 2, 3)
    s0, thw0 = skips[0]
    s0 = s0.reshape(B, *thw0, s0.shape[2]).permute(0, 4, 1, 2, 3)
    f = f + F.interpolate(s0, size=(thw0[0] * 2, thw0[1], thw0[2]), mode="trilinear")
    logits = F.conv3d(f, sd["classifier.weight"], sd["classifier.bias"])             # ref :481
    out = [logits]
    if return_embed:                                                                  # ref :493-498
        out.append(rnd_b(F.linear(rnd_f(xw.mean(dim=1)), rnd_f(sd["vision_proj.weight"]), sd["vision_proj.bias"])))
        out.append(rnd_b(F.linear(rnd_f(yw.mean(dim=1)), rnd_f(sd["audio_proj.weight"]), sd["audio_proj.bias"])))
    if return_intermediates:
        return out, inter
    if not return_embed and (return_spatial_attn or return_temporal_attn):               # ref :485-491
        return [logits] + ([p_sp] if return_spatial_attn else []) + ([p_tm] if return_temporal_attn else [])
    return out if return_embed else logits


# --------------------------------------------------------------------------------------------
# Loss.  ref: tools/train_avgaze_net.py:76-88
# --------------------------------------------------------------------------------------------

def frame_softmax(logits, temperature=2.0):
    """ref: slowfast/utils/utils.py:5-12 — softmax over H*W per (b, t)."""
    B, _, T, H, W = logits.shape
    return F.softmax(logits.reshape(B, -1, T, H * W) / temperature, dim=-1).reshape(B, -1, T, H, W)


def kldiv(pred, target):
    """ref: slowfast/models/losses.py:59-82 (target given)."""
    B, T = pred.shape[0], pred.shape[2]
    HW = pred.shape[3] * pred.shape[4]
    p = pred.reshape(B, T, -1)
    q = target.reshape(B, T, -1)
    per_frame = (p * torch.log(p + 1e-10)).sum(-1) - (p * torch.log(q + 1e-10)).sum(-1)
    return (per_frame.sum(-1) / (T * math.log(HW))).mean()


def sim_matrix(a, b, eps=1e-8):
    """ref: slowfast/utils/utils.py:15-24 — cosine similarity with norm clamp."""
    an = a / a.norm(dim=1, keepdim=True).clamp_min(eps)
    bn = b / b.norm(dim=1, keepdim=True).clamp_min(eps)
    return an @ bn.t()


def egonce(sim, temperature=0.05):
    """ref: slowfast/models/losses.py:157-170 with the eye mask on sim.device (the reference's
    `.cuda()` at :158 is the single line a CPU restatement has to change)."""
    z = sim / temperature
    diag = torch.diagonal(z)
    li = (diag - torch.logsumexp(z, dim=1)).mean()
    lj = (diag - torch.logsumexp(z, dim=0)).mean()
    return -li - lj


def kldiv_egonce(logits, v, a, labels_hm, alpha=0.05):
    """Total training loss and its two terms.  ref: tools/train_avgaze_net.py:84-88."""
    kld = kldiv(frame_softmax(logits, 2.0), labels_hm)
    nce = egonce(sim_matrix(v, a))
    return kld + alpha * nce, kld, nce


def loss_and_grads(sd, video, audio, labels_hm, alpha=0.05, loss_scale=1.0, drop_scales=None):
    """Forward + loss + autograd backward over every tensor in `sd` (fp32).  Returns
    (loss, kld, nce, logits, grads{name: tensor}).  loss_scale mirrors GradScaler (backward runs on
    loss_scale * loss, the returned gradients are unscaled); it only matters under EMULATE_BF16 with an
    fp16 BWD_DTYPE."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    logits, v, a = csts_forward(leaves, video, audio, return_embed=True, drop_scales=drop_scales)
    loss, kld, nce = kldiv_egonce(logits, v, a, labels_hm, alpha)
    names = list(leaves)
    gs = torch.autograd.grad(loss * loss_scale, [leaves[n] for n in names], allow_unused=True)
    grads = {n: g / loss_scale for n, g in zip(names, gs) if g is not None}
    return loss.detach(), kld.detach(), nce.detach(), logits.detach(), grads


# --------------------------------------------------------------------------------------------
# Metric.  ref: slowfast/utils/metrics.py:9-74 and the rescale of tools/train_avgaze_net.py:125-127
# --------------------------------------------------------------------------------------------

def minmax_rescale(preds):
    """ref: tools/train_avgaze_net.py:125-127 — per-frame (p - min) / (max - min + 1e-6)."""
    flat = preds.detach().view(preds.size()[:-2] + (preds.size(-1) * preds.size(-2),))
    mn, mx = flat.min(dim=-1, keepdim=True)[0], flat.max(dim=-1, keepdim=True)[0]
    return ((flat - mn) / (mx - mn + 1e-6)).view(preds.size())


def adaptive_f1(preds, labels_hm, labels, dataset):
    """ref: slowfast/utils/metrics.py:30-74 (the PyTorch branch), computed per threshold instead of through the two
    (thresholds, B, T, H, W) temporaries.  Returns (f1, recall, precision, threshold)."""
    import numpy as np
    if "forecast" in dataset and "aria" not in dataset:
        thresholds = np.linspace(0.01, 0.07, 31)
    elif "forecast" in dataset and "aria" in dataset:
        thresholds = np.linspace(0.0, 0.02, 21)
    else:
        thresholds = np.linspace(0, 0.02, 11)
    fixation_idx = 1 if dataset == "egteagaze" else 0
    binary_labels = (labels_hm > 0.001).float()
    tracked = torch.where(labels.reshape(-1, labels.shape[-1])[:, 2] == fixation_idx)[0]
    f1s, rcs, prs = [], [], []
    for th in thresholds:
        bp = (preds.squeeze(1) > th).float()
        tp = (bp * binary_labels).sum(dim=(2, 3)).reshape(-1).index_select(0, tracked)
        fgl = binary_labels.sum(dim=(2, 3)).reshape(-1).index_select(0, tracked)
        fgp = bp.sum(dim=(2, 3)).reshape(-1).index_select(0, tracked)
        rc, pr = (tp / (fgl + 1e-6)).mean(), (tp / (fgp + 1e-6)).mean()
        rcs.append(rc)
        prs.append(pr)
        f1s.append((2 * rc * pr) / (rc + pr + 1e-6))
    f1 = torch.stack(f1s)
    i = int(torch.argmax(f1))
    return float(f1[i]), float(rcs[i]), float(prs[i]), thresholds[i]


# --------------------------------------------------------------------------------------------
# Synthetic weights and inputs
# --------------------------------------------------------------------------------------------

def synthetic_state(shapes, seed=0, gain=1.0):
    """Deterministic weights for every (name, shape) of the reference state_dict contract
    (tests/golden/param_shapes.json).  Statistics follow the reference init
    (custom_multimodal_builder.py:304-325: trunc-normal 0.02 Linear / pos-embeds, conv
    kaiming-uniform) but biases and LayerNorm affine parameters are made non-trivial so that a
    dropped bias or swapped gamma/beta cannot hide.  `gain` scales the Linear weights (peaked
    heat-maps, SURVEY.md §7 "hard parts")."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        r = torch.randn(shape, generator=g)
        leaf = name.rsplit(".", 2)[-2] if "." in name else name
        is_norm = leaf.startswith("norm")
        if name.startswith("pos_embed"):
            t = 0.02 * r
        elif is_norm and name.endswith(".weight"):
            t = 1.0 + 0.1 * r
        elif is_norm:
            t = 0.1 * r
        elif name.endswith(".bias"):
            t = 0.02 * r
        elif len(shape) == 5:                       # conv kernels: std of kaiming-uniform(a=sqrt(5))
            fan_in = shape[1] * shape[2] * shape[3] * shape[4]
            t = r / math.sqrt(3.0 * fan_in)
        else:
            t = 0.02 * gain * r
        sd[name] = t
    return sd


def gaussian_1d(ksize=19):
    # cv2.getGaussianKernel(ksize, -1): sigma = 0.3*((ksize-1)*0.5 - 1) + 0.8
    sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    r = torch.arange(ksize, dtype=torch.float64) - (ksize - 1) / 2
    k = torch.exp(-(r * r) / (2 * sigma * sigma))
    return (k / k.sum()).float()


def synthetic_batch(B, seed=1, T=8, size=256, hm=64, ksize=19):
    """video ~ N(0,1); audio = clamp(2*N(0,1)-5, log(1e-6)); labels_hm = one 19x19 Gaussian blob
    per frame at a uniform-random centre, renormalised to sum 1
    (ref: slowfast/datasets/ego4d_avgaze_forecast.py:318-328,404-422)."""
    g = torch.Generator().manual_seed(seed)
    video = torch.randn(B, 3, T, size, size, generator=g)
    audio = (torch.randn(B, 1, T, size, size, generator=g) * 2 - 5).clamp_min(math.log(1e-6))
    k1 = gaussian_1d(ksize)
    k2 = torch.outer(k1, k1)
    labels = torch.zeros(B, T, hm, hm)
    centres = torch.randint(0, hm, (B, T, 2), generator=g)
    half = ksize // 2
    for b in range(B):
        for t in range(T):
            cy, cx = int(centres[b, t, 0]), int(centres[b, t, 1])
            y0, y1 = max(cy - half, 0), min(cy + half + 1, hm)
            x0, x1 = max(cx - half, 0), min(cx + half + 1, hm)
            labels[b, t, y0:y1, x0:x1] = k2[y0 - cy + half: y1 - cy + half, x0 - cx + half: x1 - cx + half]
            labels[b, t] /= labels[b, t].sum()
    return video, audio, labels
