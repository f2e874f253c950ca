#!/usr/bin/env python3
"""Per-kernel-family roofline table from the ncu launch list (profiles/rNN_kernel_dram.json, written by
tools/ncu_summary.py) and MEASURED_PEAKS.json.

    python tools/roofline_table.py profiles/r01_kernel_dram.json > profiles/r01_roofline_by_kernel.md

DRAM GB/s = ncu dram__bytes_read+write.sum / gpu__time_duration.sum per family (cold caches, serialised launches:
memory-bound kernels whose inputs sit in L2 in the real step are understated)."""
import json
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILIES = [
    ("gemm_tc_kernel", "tcgen05 GEMM (Linear fwd/dX/dW, attention products, fused softmax / GELU / bias-grad epilogues)"),
    ("gemm_mma_kernel", "mma.sync GEMM (skinny / odd shapes)"),
    ("dwconv_tcol_kernel", "pooling conv fwd / adjoint (+LayerNorm), T-column"),
    ("dwconv_wgrad", "pooling conv weight gradient"),
    ("dwconv_kernel", "pooling conv, generic temporal extent"),
    ("layernorm_bwd", "LayerNorm backward (+ residual add, + 16-bit operand copy)"),
    ("layernorm_fwd", "LayerNorm forward"),
    ("softmax_", "attention softmax fwd/bwd (blocks with > 256 keys, spatial fusion)"),
    ("mt_", "optimizer: grad norm + clip + AdamW + weight refresh"),
    ("cast_", "f32 -> 16-bit gradient casts"),
    ("upsample", "trilinear skip up-sampling fwd/bwd"),
    ("maxpool_", "MaxPool3d skip fwd/bwd"),
    ("im2col", "patch-embed im2col"),
    ("classifier", "classifier + stem skip fwd/bwd"),
    ("rowdot", "softmax-backward row term rowsum(dO o O)"),
    ("f1_", "adaptive-F1 metric"),
    ("colsum", "column sums (residual bias gradients)"),
    ("permute_021", "frame-pool activation transposes"),
    ("kldiv|sim_matrix|egonce|reweight|token_mean|pos_embed|add_kernel|scale_kernel", "losses, fusion glue, stem glue"),
    ("at::", "torch elementwise (zeros, cat, rand, copies)"),
]


def main():
    d = json.load(open(sys.argv[1]))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = 6555.8
    if os.path.exists(peaks_path):
        hbm = json.load(open(peaks_path)).get("hbm_gbs", hbm)
    fam = defaultdict(lambda: [0, 0.0, 0.0])
    for name, v in d.items():
        for pat, _ in FAMILIES:
            if any(name.startswith(p) or (p in name and p == "at::") for p in pat.split("|")):
                f = fam[pat]
                break
        else:
            f = fam["other"]
        f[0] += v["launches"]
        f[1] += v["ms"]
        f[2] += v["dram_bytes"]
    total = sum(v[1] for v in fam.values())
    print(f"Kernel families of one training step (ncu, {sum(v[0] for v in fam.values())} launches, {total:.2f} ms serialised / cold caches; "
          f"HBM peak {hbm:.0f} GB/s measured)\n")
    print("| family | launches | ms | share | DRAM GB | DRAM GB/s | frac of HBM peak |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    names = dict(FAMILIES)
    for pat, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        gbs = v[2] / (v[1] * 1e-3) / 1e9 if v[1] else 0.0
        print(f"| {names.get(pat, pat)} | {v[0]} | {v[1]:.2f} | {100 * v[1] / total:.1f}% | {v[2] / 1e9:.2f} | {gbs:.0f} | {gbs / hbm:.2f} |")


if __name__ == "__main__":
    main()
