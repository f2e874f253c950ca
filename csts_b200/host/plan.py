"""Static architecture plan of CSTS derived from the cfg (no tensors).

Follows the constructor arithmetic of ``custom_multimodal_builder.py:107-300``: dimension / head
multipliers, q-pool strides and the adaptive kv stride for the 16 video blocks, the hard-coded audio
encoder (:184-191), fusion blocks (:232-270) and decoder (:272-300).
"""
from dataclasses import dataclass
from typing import Optional, Tuple

import torch


@dataclass(frozen=True)
class BlockSpec:
    name: str                       # module path, e.g. "blocks.3" / "decode_block2" / "spatial_fusion"
    kind: str                       # "enc" | "dec" | "spatial" | "temporal"
    dim: int
    dim_out: int
    heads: int
    stride_q: Optional[Tuple[int, int, int]]    # None: q is not pooled / up-sampled
    stride_kv: Optional[Tuple[int, int, int]]   # None: k, v are not pooled (fusion blocks)
    drop_path: float

    @property
    def head_dim(self):
        return self.dim // self.heads

    @property
    def hidden(self):
        # encoder / fusion MLP: 4*dim (attention.py:205); decoder: 4*dim_out (attention.py:444)
        return 4 * (self.dim_out if self.kind == "dec" else self.dim)

    def q_grid(self, thw):
        if self.stride_q is None:
            return tuple(thw)
        if self.kind == "dec":
            return tuple(n * s for n, s in zip(thw, self.stride_q))
        return tuple((n - 1) // s + 1 for n, s in zip(thw, self.stride_q))

    def kv_grid(self, thw):
        if self.stride_kv is None:
            return tuple(thw)
        return tuple((n - 1) // s + 1 for n, s in zip(thw, self.stride_kv))


def round_width(width, multiplier, min_width=1, divisor=1):
    """slowfast/models/utils.py:8-21."""
    if not multiplier:
        return width
    width *= multiplier
    min_width = min_width or divisor
    out = max(min_width, int(width + divisor / 2) // divisor * divisor)
    if out < 0.9 * width:
        out += divisor
    return int(out)


def build_plan(cfg):
    mv = cfg.MVIT
    depth = mv.DEPTH
    assert mv.MODE == "conv" and not mv.POOL_FIRST, "CSTS configs use conv pooling after projection"
    dim_mul, head_mul = [1.0] * (depth + 1), [1.0] * (depth + 1)
    for i, m in mv.DIM_MUL:
        dim_mul[i] = m
    for i, m in mv.HEAD_MUL:
        head_mul[i] = m
    stride_q = {int(r[0]): tuple(int(s) for s in r[1:]) for r in mv.POOL_Q_STRIDE}
    # adaptive kv stride (custom_multimodal_builder.py:129-136)
    if mv.POOL_KV_STRIDE_ADAPTIVE is not None:
        skv = [int(s) for s in mv.POOL_KV_STRIDE_ADAPTIVE]
        kv = {}
        for i in range(depth):
            if i in stride_q:
                skv = [max(skv[d] // stride_q[i][d], 1) for d in range(3)]
            kv[i] = tuple(skv)
    else:
        kv = {int(r[0]): tuple(int(s) for s in r[1:]) for r in (mv.POOL_KV_STRIDE or [])}
    assert mv.POOL_KVQ_KERNEL is None or list(mv.POOL_KVQ_KERNEL) == [3, 3, 3], "kernels are specialised for 3x3x3 pooling"
    dpr = [x.item() for x in torch.linspace(0, mv.DROPPATH_RATE, depth)]        # :90
    specs = []
    heads, dim = mv.NUM_HEADS, mv.EMBED_DIM
    for i in range(depth):
        heads = round_width(heads, head_mul[i])
        dim = round_width(dim, dim_mul[i], divisor=heads)
        dim_out = round_width(dim, dim_mul[i + 1], divisor=round_width(heads, head_mul[i + 1]))
        specs.append(BlockSpec(f"blocks.{i}", "enc", dim, dim_out, heads, stride_q.get(i), kv.get(i), dpr[i]))
    a_dim, a_out, a_heads = [96, 192, 384, 768], [192, 384, 768, 768], [1, 2, 4, 8]
    a_sq = [None, (1, 2, 2), (1, 2, 2), (1, 2, 2)]
    a_skv = [(1, 8, 8), (1, 4, 4), (1, 2, 2), (1, 1, 1)]
    for i in range(4):
        specs.append(BlockSpec(f"blocks_audio.{i}", "enc", a_dim[i], a_out[i], a_heads[i], a_sq[i], a_skv[i], 0.0))
    tok = specs[depth - 1].dim_out
    specs.append(BlockSpec("temporal_fusion", "temporal", tok, tok, heads, None, None, 0.0))
    specs.append(BlockSpec("spatial_fusion", "spatial", tok, tok, heads, None, None, 0.0))
    d_in, d_out, d_heads = [768, 768, 384, 192], [768, 384, 192, 96], [8, 4, 4, 2]
    d_sq = [(1, 2, 2), (1, 2, 2), (1, 2, 2), (2, 1, 1)]
    d_skv = [(1, 2, 2), (1, 4, 4), (1, 8, 8), (1, 16, 16)]
    for i in range(4):
        specs.append(BlockSpec(f"decode_block{i + 1}", "dec", d_in[i], d_out[i], d_heads[i], d_sq[i], d_skv[i], 0.0))
    return specs
