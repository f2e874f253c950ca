"""Attention-pooling microbenchmark at every distinct pool shape of CSTS (SURVEY.md App. A.2, B=8):
depthwise conv (+LayerNorm) forward, data gradient, weight gradient — achieved GB/s of algorithmic
bytes (input once + output once, bf16) against the measured HBM copy bandwidth.  CUDA-graph timed."""
import json
import sys

import torch

sys.path.insert(0, ".")
from csts_b200 import kernels as K  # noqa: E402

B = 8
# (name, heads, d, thw_in, stride, transposed)
SHAPES = [
    ("blocks.0 kv", 1, 96, (4, 64, 64), (1, 8, 8), False), ("blocks.1 q", 2, 96, (4, 64, 64), (1, 2, 2), False),
    ("blocks.1 kv", 2, 96, (4, 64, 64), (1, 4, 4), False), ("blocks.2 kv", 2, 96, (4, 32, 32), (1, 4, 4), False),
    ("blocks.3 qkv", 4, 96, (4, 32, 32), (1, 2, 2), False), ("blocks.4-13 kv", 4, 96, (4, 16, 16), (1, 2, 2), False),
    ("blocks.14 q", 8, 96, (4, 16, 16), (1, 2, 2), False), ("blocks.14 kv", 8, 96, (4, 16, 16), (1, 1, 1), False),
    ("blocks.15 kv", 8, 96, (4, 8, 8), (1, 1, 1), False), ("dec1 kv", 8, 96, (4, 8, 8), (1, 2, 2), False),
    ("dec2 kv", 4, 192, (4, 16, 16), (1, 4, 4), False), ("dec3 kv", 4, 96, (4, 32, 32), (1, 8, 8), False),
    ("dec4 kv", 2, 96, (4, 64, 64), (1, 16, 16), False), ("dec1 q", 8, 96, (4, 8, 8), (1, 2, 2), True),
    ("dec2 q", 4, 192, (4, 16, 16), (1, 2, 2), True), ("dec3 q", 4, 96, (4, 32, 32), (1, 2, 2), True),
    ("dec4 q", 2, 96, (4, 64, 64), (2, 1, 1), True),
]


def graph_time(fn, reps=10):
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
    return s.elapsed_time(e) / reps * 1e3   # us


def main():
    dev = "cuda"
    peak = 6555.8
    for name, h, d, thw, stride, tr in SHAPES:
        N = thw[0] * thw[1] * thw[2]
        Cn = h * d
        qkv = torch.randn(B, N, 3, h, d, device=dev).to(torch.bfloat16)
        w = torch.randn(d, 1, 3, 3, 3, device=dev) * 0.2
        gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
        qs = (N * 3 * Cn, d, 3 * Cn)
        out, pre, mean, rstd, thw_o = K.dwconv(qkv, qs, Cn, B, h, d, thw, stride, w, transposed=tr, norm=(gamma, beta))
        Lo = thw_o[0] * thw_o[1] * thw_o[2]
        dense = (h * Lo * d, Lo * d, d)
        du = torch.randn(B, h, Lo, d, device=dev).to(torch.bfloat16)
        dqkv = torch.zeros_like(qkv)
        dw = torch.zeros_like(w)
        t_f = graph_time(lambda: K.dwconv(qkv, qs, Cn, B, h, d, thw, stride, w, transposed=tr, norm=(gamma, beta)))
        t_b = graph_time(lambda: K.dwconv(du, dense, 0, B, h, d, thw_o, stride, w, transposed=not tr, out=dqkv, out_strides=qs,
                                          out_off=Cn, thw_out=thw))
        if tr:
            t_w = graph_time(lambda: K.dwconv_wgrad(qkv, qs, Cn, thw, du, dense, 0, thw_o, B, h, d, stride, dw))
        else:
            t_w = graph_time(lambda: K.dwconv_wgrad(du, dense, 0, thw_o, qkv, qs, Cn, thw, B, h, d, stride, dw))
        b_in, b_out = B * h * N * d * 2, B * h * Lo * d * 2
        row = dict(shape=name, in_MB=b_in / 1e6, out_MB=b_out / 1e6, fwd_us=t_f, bwd_data_us=t_b, wgrad_us=t_w,
                   fwd_gbs=(b_in + 2 * b_out) / t_f / 1e3, bwd_gbs=(b_in + b_out) / t_b / 1e3, wgrad_gbs=(b_in + b_out) / t_w / 1e3)
        row["fwd_frac"] = row["fwd_gbs"] / peak
        print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in row.items()}), flush=True)


if __name__ == "__main__":
    main()
