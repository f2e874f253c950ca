"""CSTS — audio-visual gaze-forecasting network on the libcsts_b200 kernels.

Drop-in for ``slowfast/models/custom_multimodal_builder.py:20-498``: same constructor argument
(`cfg`), same ``forward(x, y, return_embed, return_spatial_attn, return_temporal_attn)`` contract,
same module tree and parameter names/shapes/order (so reference checkpoints, the reference
optimizer's parameter grouping and DDP work unchanged), same initialisation (identical tensors
under the same ``torch.manual_seed``).  The ``nn.Linear`` / ``nn.Conv3d`` / ``nn.LayerNorm`` members
are *parameter holders only*: their ``forward`` is never called — every operation runs in
hand-written CUDA through ``csts_b200.kernels``; there is no PyTorch or CPU fallback.
"""
import math
import os
from functools import partial

import torch
import torch.nn as nn
from torch.nn.init import trunc_normal_

from .. import kernels as K
from .block import BlockFn, block_forward, block_param_names
from .build import MODEL_REGISTRY
from .grad_arena import GradArena, grad_slot
from .plan import build_plan
from .weights import WeightCache, precision_of


# ---------------------------------------------------------------------------------------------
# parameter holders (module tree mirrors the reference: attention.py, av_attention.py, common.py,
# stem_helper.py)
# ---------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("parameter holder: CSTS runs through csts_b200 kernels, not nn.Module.forward")


class PatchEmbed(_Holder):
    def __init__(self, dim_in, dim_out, kernel, stride, padding):
        super().__init__()
        self.proj = nn.Conv3d(dim_in, dim_out, kernel_size=tuple(kernel), stride=tuple(stride), padding=tuple(padding))


class Mlp(_Holder):
    def __init__(self, dim, hidden, dim_out):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim_out)


class Attention(_Holder):
    def __init__(self, spec, qkv_bias):
        super().__init__()
        dim, d = spec.dim, spec.head_dim
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        if spec.stride_q is not None:
            if spec.kind == "dec":
                op = tuple(0 if s == 1 else s - 1 for s in spec.stride_q)
                self.upsample_q = nn.ConvTranspose3d(d, d, (3, 3, 3), stride=spec.stride_q, padding=(1, 1, 1), output_padding=op,
                                                     groups=d, bias=False)
            else:
                self.pool_q = nn.Conv3d(d, d, (3, 3, 3), stride=spec.stride_q, padding=(1, 1, 1), groups=d, bias=False)
            self.norm_q = nn.LayerNorm(d)
        if spec.stride_kv is not None:
            self.pool_k = nn.Conv3d(d, d, (3, 3, 3), stride=spec.stride_kv, padding=(1, 1, 1), groups=d, bias=False)
            self.norm_k = nn.LayerNorm(d)
            self.pool_v = nn.Conv3d(d, d, (3, 3, 3), stride=spec.stride_kv, padding=(1, 1, 1), groups=d, bias=False)
            self.norm_v = nn.LayerNorm(d)


class Block(_Holder):
    """MultiScaleBlock / MultiScaleDecoderBlock / SpatialBlock / TemporalBlock parameter set."""

    def __init__(self, spec, qkv_bias, norm_layer):
        super().__init__()
        self.spec = spec
        self.dim, self.dim_out = spec.dim, spec.dim_out
        self.norm1 = norm_layer(spec.dim)
        self.attn = Attention(spec, qkv_bias)
        self.norm2 = norm_layer(spec.dim)
        self.mlp = Mlp(spec.dim, spec.hidden, spec.dim_out)
        if spec.dim != spec.dim_out:
            self.proj = nn.Linear(spec.dim, spec.dim_out)
        self._names = block_param_names(spec)

    def tensors(self):
        out = []
        for n in self._names:
            obj = self
            for part in n.split("."):
                obj = getattr(obj, part)
            out.append(obj)
        return out


# ---------------------------------------------------------------------------------------------
# autograd nodes of the model-level glue
# ---------------------------------------------------------------------------------------------
class PatchEmbedFn(torch.autograd.Function):
    """Conv3d k(3,7,7) s(2,4,4) p(1,3,3) as im2col + tcgen05 GEMM, bias and the separable position
    embedding fused in the epilogue.  ref: stem_helper.py:35-38, custom_multimodal_builder.py:345,362-370."""

    @staticmethod
    def forward(ctx, wc, x, weight, bias, pos_spatial, pos_temporal):
        B = x.shape[0]
        kdim = weight[0].numel()
        kp = (kdim + 7) // 8 * 8
        patches = K.im2col_patch(x.contiguous(), kp, dtype=wc.act)
        pos = K.pos_embed(pos_spatial, pos_temporal)
        n_tok = pos.shape[0]
        tok = K.gemm(patches, wc.w_padded(weight, kp), M=B * n_tok, N=weight.shape[0], K=kp, bias=bias, residual=pos, res_mod=n_tok,
                     out_dtype=torch.float32)
        ctx.save_for_backward(patches)
        ctx.wc = wc
        ctx.dims = (B, n_tok, pos_temporal.shape[1], pos_spatial.shape[1], weight.shape, kp)
        return tok.view(B, n_tok, weight.shape[0])

    @staticmethod
    def backward(ctx, dtok):
        (patches,) = ctx.saved_tensors
        B, n_tok, T, HW, wshape, kp = ctx.dims
        Cn = wshape[0]
        dtok = dtok.contiguous()
        g = K.cast16(dtok.view(B * n_tok, Cn), ctx.wc.grad)
        M = B * n_tok
        dwp = K.gemm(g, patches, M=Cn, N=kp, K=M, a_kmajor=False, b_kmajor=False, lda=Cn, ldb=kp, out_dtype=torch.float32,
                     split_k=max(1, min(64, M // 2048)))
        kdim = wshape[1] * wshape[2] * wshape[3] * wshape[4]
        dw = dwp[:, :kdim].reshape(wshape)
        db = K.colsum(dtok, M, Cn)
        dsp, dtm = K.pos_embed_bwd(dtok, B, T, HW, Cn)
        return None, None, dw, db, dsp, dtm


class FramePoolFn(torch.autograd.Function):
    """Dense Conv3d(C, C, (1,8,8)) over a (B, T*64, C) token map = skinny split-K GEMM
    (M = B*T, K = 64*C, N = C).  ref: custom_multimodal_builder.py:227-229, :420-421, :442-445.

    The contraction index is ordered (c, hw) — the native layout of the (O, C, 1, 8, 8) parameter — by
    transposing the small activation (1.5 M elements) instead of the 37.7 M-element weight: the weight's
    operand copy is then the plain one (refreshed by the fused optimizer step) and dW needs no re-layout."""

    @staticmethod
    def forward(ctx, wc, tok, weight, bias):
        B, N, Cn = tok.shape
        T = N // 64
        a = K.permute_021(tok.contiguous(), B * T, 64, Cn, wc.act)          # (B*T, hw, c) f32 -> (B*T, c, hw) 16-bit
        Kd = 64 * Cn
        # Split over the 49152-deep contraction WITHOUT atomics: the splits are a batch dimension (operand batch stride = one
        # k-range), every split writes its own (B*T, O) partial, and a column sum adds them — in a fixed order — onto the bias.
        # The forward pass therefore holds no order-dependent f32 sum: two evaluations of one batch are bit-identical.
        O = weight.shape[0]
        S = 24 if Kd % (24 * 64) == 0 else 1
        Kc = Kd // S
        part = torch.empty((S, B * T, O), dtype=torch.float32, device=tok.device)
        K.gemm(a.view(B * T, Kd), wc.w(weight), M=B * T, N=O, K=Kc, lda=Kd, ldb=Kd, out=part, ldc=O, batch=(S, 1), sA=(Kc, 0), sB=(Kc, 0),
               sC=(B * T * O, 0))
        out = bias.detach().repeat(B * T)
        K.colsum(part.view(S, B * T * O), S, B * T * O, out=out)
        ctx.save_for_backward(a)
        ctx.wc, ctx.weight, ctx.bias, ctx.dims = wc, weight, bias, (B, N, Cn)
        return out.view(B, T, weight.shape[0])

    @staticmethod
    def backward(ctx, dout):
        (a,) = ctx.saved_tensors
        weight = ctx.weight
        B, N, Cn = ctx.dims
        T, O, Kd = N // 64, weight.shape[0], 64 * Cn
        dout = dout.contiguous().view(B * T, O)
        g = K.cast16(dout, ctx.wc.grad)
        dtok_t = K.gemm(g, ctx.wc.w(weight), M=B * T, N=Kd, K=O, b_kmajor=False, ldb=Kd, out_dtype=torch.float32)   # (B*T, c, hw)
        dtok = K.permute_021(dtok_t, B * T, Cn, 64, torch.float32)                                                  # -> (B*T, hw, c)
        slot = grad_slot(ctx.wc, weight)          # 151 MB: written straight into the gradient arena
        sync = getattr(ctx.wc, "grad_sync", None)
        if sync is not None and slot is not None and id(weight) in sync.factored:
            # data parallel: the ranks exchange the two thin factors of this gradient instead of the 151 MB product
            # (host/distributed.py::OverlappedGradSync.factored_wgrad); the bias gradient takes the ordinary all-reduce
            db = K.colsum(dout, B * T, O)
            dw = sync.factored_wgrad(weight, g, a.view(B * T, Kd), slot.view(O, Kd))
            return None, dtok.view(B, N, Cn), dw.view(weight.shape), db
        bslot = grad_slot(ctx.wc, ctx.bias)
        if slot is not None and bslot is not None:
            # both gradients land in arena slots nobody reads before the optimizer: the 151 MB product (and the bias gradient
            # fused into it) leaves the dX chain for the weight-gradient stream, like the blocks' weight gradients
            fork = ctx.wc.backward_fork()

            def wgrad():
                bslot.zero_()
                return K.gemm(g, a.view(B * T, Kd), M=O, N=Kd, K=B * T, a_kmajor=False, b_kmajor=False, lda=O, ldb=Kd,
                              out_dtype=torch.float32, out=slot.view(O, Kd), rowsum=bslot)
            dw = fork.run(wgrad, g, a)
            if not getattr(ctx.wc, "defer_join", False):
                fork.join()
            return None, dtok.view(B, N, Cn), dw.view(weight.shape), bslot
        db = torch.zeros(O, dtype=torch.float32, device=dout.device)
        dw = K.gemm(g, a.view(B * T, Kd), M=O, N=Kd, K=B * T, a_kmajor=False, b_kmajor=False, lda=O, ldb=Kd, out_dtype=torch.float32,
                    out=None if slot is None else slot.view(O, Kd), rowsum=db)
        return None, dtok.view(B, N, Cn), dw.view(weight.shape), db


class ReweightFn(torch.autograd.Function):
    """x * w[:, t] broadcast over the 64 positions of frame t; `w_off` selects the visual (0) or audio
    (T*C) half of the temporal-fusion output.  ref: custom_multimodal_builder.py:454-461."""

    @staticmethod
    def forward(ctx, x, av, w_off, T):
        B, N, Cn = x.shape
        x, av = x.contiguous(), av.contiguous()
        out = K.reweight_fwd(x, av, w_off, av.shape[1] * Cn, B, T, N // T, Cn)
        ctx.save_for_backward(x, av)
        ctx.w_off, ctx.T = w_off, T
        return out

    @staticmethod
    def backward(ctx, dout):
        x, av = ctx.saved_tensors
        B, N, Cn = x.shape
        dav = torch.zeros_like(av)
        dx = K.reweight_bwd(dout.contiguous(), x, av, ctx.w_off, av.shape[1] * Cn, dav, B, ctx.T, N // ctx.T, Cn)
        return dx, dav, None, None


class MeanProjFn(torch.autograd.Function):
    """mean over tokens followed by Linear(768, 256).  ref: custom_multimodal_builder.py:493-496."""

    @staticmethod
    def forward(ctx, wc, x, weight, bias):
        B, N, Cn = x.shape
        m = K.token_mean_fwd(x.contiguous(), B, N, Cn, dtype=wc.act)
        out = K.gemm(m, wc.w(weight), M=B, N=weight.shape[0], K=Cn, bias=bias, out_dtype=torch.float32)
        ctx.save_for_backward(m)
        ctx.wc, ctx.weight, ctx.dims = wc, weight, (B, N, Cn)
        return out

    @staticmethod
    def backward(ctx, dout):
        (m,) = ctx.saved_tensors
        B, N, Cn = ctx.dims
        O = ctx.weight.shape[0]
        dout = dout.contiguous()
        g = K.cast16(dout, ctx.wc.grad)
        dm = K.gemm(g, ctx.wc.w(ctx.weight), M=B, N=Cn, K=O, b_kmajor=False, ldb=Cn, out_dtype=torch.float32)
        dx = K.token_mean_bwd(dm, B, N, Cn)
        dw = K.gemm(g, m, M=O, N=Cn, K=B, a_kmajor=False, b_kmajor=False, lda=O, ldb=Cn, out_dtype=torch.float32)
        db = K.colsum(dout, B, O)
        return None, dx, dw, db


class HeadFn(torch.autograd.Function):
    """classifier(feat + trilinear_T(stem)) -> logits (B,1,2T,H,W).  ref: custom_multimodal_builder.py:476-481."""

    @staticmethod
    def forward(ctx, feat, stem, weight, bias, thw):
        B, _, Cn = feat.shape
        Ti, S = thw[0], thw[1] * thw[2]
        feat, stem = feat.contiguous(), stem.contiguous()
        logits = K.classifier_fwd(feat, stem, weight, bias, B, Ti, S, Cn)
        ctx.save_for_backward(feat, stem, weight)
        ctx.dims = (B, Ti, S, Cn)
        return logits.view(B, 1, 2 * Ti, thw[1], thw[2])

    @staticmethod
    def backward(ctx, dlogits):
        feat, stem, weight = ctx.saved_tensors
        B, Ti, S, Cn = ctx.dims
        dfeat, dstem, dw, db = K.classifier_bwd(dlogits.contiguous(), feat, stem, weight, B, Ti, S, Cn)
        return dfeat, dstem, dw.view_as(weight), db, None


class AddFn(torch.autograd.Function):
    """Decoder skip connection a + b on the f32 residual stream.  ref: custom_multimodal_builder.py:467-473."""

    @staticmethod
    def forward(ctx, a, b):
        return K.add_f32(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, g):
        return g, g


# ---------------------------------------------------------------------------------------------
@MODEL_REGISTRY.register()
class CSTS(nn.Module):
    """Multiscale Vision Transformer with audio-visual fusion (see module docstring)."""

    def __init__(self, cfg):
        super().__init__()
        assert cfg.DATA.TRAIN_CROP_SIZE == cfg.DATA.TEST_CROP_SIZE
        self.cfg = cfg
        mv = cfg.MVIT
        assert not mv.CLS_EMBED_ON and mv.SEP_POS_EMBED and not mv.PATCH_2D and not mv.NORM_STEM, \
            "csts_b200 implements the configuration of configs/{Ego4D,Aria}/CSTS_*.yaml"
        assert mv.DROPOUT_RATE == 0.0
        assert mv.NORM == "layernorm"
        assert list(mv.PATCH_KERNEL) == [3, 7, 7] and list(mv.PATCH_STRIDE) == [2, 4, 4] and list(mv.PATCH_PADDING) == [1, 3, 3], \
            "the patch-embed kernel is specialised for k(3,7,7) s(2,4,4) p(1,3,3)"
        self.spatial_audio_attn = mv.SPATIAL_AUDIO_ATTN
        norm_layer = partial(nn.LayerNorm, eps=1e-6)
        embed_dim = mv.EMBED_DIM
        self.patch_stride = list(mv.PATCH_STRIDE)
        size, frames = cfg.DATA.TRAIN_CROP_SIZE, cfg.DATA.NUM_FRAMES
        self.patch_dims = [frames // self.patch_stride[0], size // self.patch_stride[1], size // self.patch_stride[2]]
        specs = build_plan(cfg)
        self.specs = {s.name: s for s in specs}
        depth = mv.DEPTH

        self.patch_embed = PatchEmbed(cfg.DATA.INPUT_CHANNEL_NUM[0], embed_dim, mv.PATCH_KERNEL, mv.PATCH_STRIDE, mv.PATCH_PADDING)
        self.patch_embed_audio = PatchEmbed(1, embed_dim, mv.PATCH_KERNEL, mv.PATCH_STRIDE, mv.PATCH_PADDING)
        hw = self.patch_dims[1] * self.patch_dims[2]
        self.pos_embed_spatial = nn.Parameter(torch.zeros(1, hw, embed_dim))
        self.pos_embed_temporal = nn.Parameter(torch.zeros(1, self.patch_dims[0], embed_dim))
        self.pos_embed_spatial_audio = nn.Parameter(torch.zeros(1, hw, embed_dim))
        self.pos_embed_temporal_audio = nn.Parameter(torch.zeros(1, self.patch_dims[0], embed_dim))

        self.blocks = nn.ModuleList([Block(s, mv.QKV_BIAS, norm_layer) for s in specs[:depth]])
        self.blocks_audio = nn.ModuleList([Block(s, mv.QKV_BIAS, norm_layer) for s in specs[depth:depth + 4]])
        token_dim = specs[depth - 1].dim_out
        if "nce" in cfg.MODEL.LOSS_FUNC:
            self.vision_proj = nn.Linear(token_dim, 256)
            self.audio_proj = nn.Linear(token_dim, 256)
        self.vision_pool = nn.Conv3d(token_dim, token_dim, kernel_size=(1, 8, 8), stride=1)
        self.audio_pool = nn.Conv3d(token_dim, token_dim, kernel_size=(1, 8, 8), stride=1)
        self.audio_pool2 = nn.Conv3d(token_dim, token_dim, kernel_size=(1, 8, 8), stride=1)
        self.temporal_fusion = Block(self.specs["temporal_fusion"], mv.QKV_BIAS, norm_layer)
        self.spatial_fusion = Block(self.specs["spatial_fusion"], mv.QKV_BIAS, norm_layer)
        for i in range(1, 5):
            setattr(self, f"decode_block{i}", Block(self.specs[f"decode_block{i}"], mv.QKV_BIAS, norm_layer))
        self.classifier = nn.Conv3d(96, 1, kernel_size=1)

        trunc_normal_(self.pos_embed_spatial, std=0.02)
        trunc_normal_(self.pos_embed_temporal, std=0.02)
        trunc_normal_(self.pos_embed_spatial_audio, std=0.02)
        trunc_normal_(self.pos_embed_temporal_audio, std=0.02)
        self.apply(self._init_weights)
        self._wc = WeightCache(precision_of(cfg))
        self._wc.act_checkpoint = bool(cfg.MODEL.ACT_CHECKPOINT)       # encoder blocks recompute their forward in backward
        self._dp_site, self._dp_keep, self._dp_scales = None, None, None
        self._last_out = None

    @staticmethod
    def _init_weights(m):
        # custom_multimodal_builder.py:318-325
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        if self.cfg.MVIT.ZERO_DECAY_POS_CLS:
            return {"pos_embed_spatial", "pos_embed_temporal", "pos_embed_class"}
        return {}

    # -----------------------------------------------------------------------------------------
    def _run_block(self, blk, x, thw, extra=None):
        spec = blk.spec
        dp_scale = None
        if self.training and spec.drop_path > 0.0:
            dp_scale = self._dp_scales[self._dp_site[id(blk)]]
        # x produced by another block (directly, or through a decoder skip add): that block's backward starts from a
        # 16-bit copy of this block's dx, written by this block's closing LayerNorm backward (block.py "hand")
        hand = None
        last = self._last_out
        if last is not None and last[0] is x and torch.is_grad_enabled() and self.training:
            hand = {"dp": last[1]}
        meta = (spec, self._wc, tuple(thw), dp_scale, blk._names, extra, hand)
        y = BlockFn.apply(meta, x, *blk.tensors())
        if isinstance(y, tuple):
            self._last_out = None
        else:
            self._last_out = (y, dp_scale)
        return y, spec.q_grid(thw)

    def _draw_drop_path(self, batch, device):
        """DropPath (common.py:46-59): per-sample keep mask scaled by 1/keep_prob, floor(keep + U[0,1)) / keep.
        A block applies it twice with independent draws — to the attention branch and to the MLP branch
        (attention.py:242 and :247) — so every block gets two rows.  The masks of all blocks of the step are drawn in
        one shot (three launches instead of four per DropPath call)."""
        if self._dp_site is None:
            blks = [m for m in self.modules() if isinstance(m, Block) and m.spec.drop_path > 0.0]
            self._dp_site = {id(b): i for i, b in enumerate(blks)}
            self._dp_keep = torch.tensor([[[1.0 - b.spec.drop_path]] for b in blks], dtype=torch.float32)
        if not self._dp_site:
            return
        if self._dp_keep.device != device:
            self._dp_keep = self._dp_keep.to(device)
        u = torch.rand(len(self._dp_site), 2, batch, dtype=torch.float32, device=device)
        self._dp_scales = torch.floor(u.add_(self._dp_keep)).div_(self._dp_keep)        # (blocks, 2 branches, B)

    def _ensure_arena(self):
        """Gradient arena of the replica (host/grad_arena.py), (re)built when the parameters moved."""
        wc = self._wc
        params = list(self.parameters())
        if wc.arena is None or not wc.arena.matches(params):
            wc.arena = GradArena(params) if os.environ.get("CSTS_GRAD_ARENA", "1") == "1" else None

    def factored_grad_params(self):
        """Parameters whose gradient is a thin product (B*T rows) that a data-parallel exchange can average from its
        factors instead of all-reducing the product: the three frame-pool kernels (host/distributed.py)."""
        return [self.vision_pool.weight, self.audio_pool.weight, self.audio_pool2.weight]

    def forward(self, x, y, return_embed=False, return_spatial_attn=False, return_temporal_attn=False):
        video = x[0] if isinstance(x, (list, tuple)) else x
        if not video.is_cuda:
            raise RuntimeError("csts_b200.CSTS runs on CUDA (sm_100a) only; there is no CPU path")
        video, audio = video.float(), y.float()
        wc = self._wc
        if self.training:
            self._draw_drop_path(video.shape[0], video.device)
        if self.training and torch.is_grad_enabled():
            wc.begin_training_step()
            self._ensure_arena()
        else:
            wc.begin_inference()
        # The audio encoder (stem + 4 blocks) is independent of the video encoder until the fusion blocks
        # (custom_multimodal_builder.py:386-411 interleaves them only textually): it runs on a second stream, so its
        # kernels share the SMs with the video encoder's many sub-wave launches.  autograd replays each node on the stream
        # of its forward, so the same overlap holds in backward; inside a captured step the two become graph branches.
        thw = tuple(self.patch_dims)
        thw_a = thw
        wc.note_main_stream()
        side = wc.audio_stream()
        main = torch.cuda.current_stream() if side is not None else None

        self._last_out = None

        def audio_encoder():
            y = PatchEmbedFn.apply(wc, audio, self.patch_embed_audio.proj.weight, self.patch_embed_audio.proj.bias,
                                   self.pos_embed_spatial_audio, self.pos_embed_temporal_audio)
            t = thw_a
            for blk in self.blocks_audio:
                y, t = self._run_block(blk, y, t)
            return y, t

        if side is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                y, thw_a = audio_encoder()
        x = PatchEmbedFn.apply(wc, video, self.patch_embed.proj.weight, self.patch_embed.proj.bias,
                               self.pos_embed_spatial, self.pos_embed_temporal)
        B = x.shape[0]
        skips = [(x, thw)]
        for i, blk in enumerate(self.blocks):
            x, thw = self._run_block(blk, x, thw)
            if i in (0, 2, 13):
                skips.append((x, thw))
                self._last_out = None          # this output has a second consumer (a decoder skip): its gradient is a sum
        if side is not None:
            main.wait_stream(side)
            y.record_stream(main)          # allocated on the second stream, consumed (and saved for backward) on this one
        else:
            y, thw_a = audio_encoder()
        # spatial fusion (custom_multimodal_builder.py:414-432)
        n_vis = x.shape[1]
        y_sp = FramePoolFn.apply(wc, y, self.audio_pool.weight, self.audio_pool.bias)
        sp_extra = {"want": "audio_rows"} if self.spatial_audio_attn else ({"want": "attn"} if return_spatial_attn else None)
        av, _ = self._run_block(self.spatial_fusion, torch.cat([x, y_sp], dim=1), thw, sp_extra)
        x_tin = x
        if self.spatial_audio_attn:
            # audio-attention re-weighting of the temporal-fusion input (av_attention.py:360-370 and :438-440): min-max
            # rescale of the audio token's attention over its frame, mean over heads.  Optional path (no shipped YAML
            # enables it): the handful of small elementwise ops below run in torch, under autograd.
            av, a_rows = av                                              # (B, heads, T, HW) rows of P, differentiable
            amax = a_rows.max(dim=-1, keepdim=True)[0]
            amin = a_rows.min(dim=-1, keepdim=True)[0]
            w_a = ((a_rows - amin) / (amax - amin + 1e-8)).mean(dim=1).reshape(x.shape[0], -1, 1)
            x_tin = x * w_a
        x_sp = av[:, :n_vis]
        # temporal fusion (:435-451)
        x_t = FramePoolFn.apply(wc, x_tin, self.vision_pool.weight, self.vision_pool.bias)
        y_t = FramePoolFn.apply(wc, y, self.audio_pool2.weight, self.audio_pool2.bias)
        tm_extra = {"want": "attn"} if return_temporal_attn else None
        av_t, _ = self._run_block(self.temporal_fusion, torch.cat([x_t, y_t], dim=1), (2, 2, 2), tm_extra)
        # re-weight (:454-461)
        T = thw[0]
        Cn = x.shape[2]
        xw = ReweightFn.apply(x_sp, av_t, 0, T)
        # decoder (:465-479)
        f = xw
        for i in range(4):
            f, thw = self._run_block(getattr(self, f"decode_block{i + 1}"), f, thw)
            if i < 3:
                prev = self._last_out
                f = AddFn.apply(f, skips[3 - i][0])
                if prev is not None:
                    self._last_out = (f, prev[1])          # the add passes the gradient through unchanged
        self._last_out = None
        stem, thw0 = skips[0]
        logits = HeadFn.apply(f, stem, self.classifier.weight, self.classifier.bias, thw0)
        if not return_embed and (return_spatial_attn or return_temporal_attn):           # :485-491, visualisation outputs
            out = [logits]
            if return_spatial_attn:
                if self.spatial_audio_attn:
                    raise ValueError("return_spatial_attn is undefined under MVIT.SPATIAL_AUDIO_ATTN (as in the reference, "
                                     "custom_multimodal_builder.py:425-430)")
                out.append(sp_extra["attn"])
            if return_temporal_attn:
                out.append(tm_extra["attn"])
            return out
        if not return_embed:
            return logits
        yw = ReweightFn.apply(y, av_t, T * Cn, T)
        v = MeanProjFn.apply(wc, xw, self.vision_proj.weight, self.vision_proj.bias)
        a = MeanProjFn.apply(wc, yw, self.audio_proj.weight, self.audio_proj.bias)
        return [logits, v, a]
