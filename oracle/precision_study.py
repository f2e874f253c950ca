"""Where does the bf16 path's gradient deviation from fp32 come from?  CPU-only study on the oracle
(TEST INFRASTRUCTURE): run the bf16-emulating oracle with individual rounding points disabled and report
the global relative L2 error of the full gradient against the fp32 oracle (B=2, synthetic_state seed 0).

    python oracle/precision_study.py

Round-1 result (8 threads, ~4 min): all roundings 3.04e-2 | forward roundings only 2.99e-2 | backward
roundings only 1.63e-2 | no single tensor class dominates (each moves the figure by < 0.3e-2).  The
deviation is therefore the forward activation precision acting through the loss gradient (which is a small
difference of log-probabilities), not the rounding of gradients."""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import csts_oracle as O  # noqa: E402

TAGS = ["pre", "pooled", "w", "qkv", "dS", "o", "g1", "xn", "dZ", "h", "g2", "wproj", "gproj", "patchx", "patchw", "patchg"]


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    shapes = json.load(open(os.path.join(os.path.dirname(HERE), "tests", "golden", "param_shapes.json")))
    sd = O.synthetic_state(shapes, seed=0)
    video, audio, hm = O.synthetic_batch(2, seed=1)
    _, _, _, flog, fg = O.loss_and_grads(sd, video, audio, hm)
    den = sum(g.pow(2).sum().item() for g in fg.values())

    def run(skip):
        O.EMULATE_BF16, O.SKIP = True, set(skip)
        try:
            _, _, _, elog, eg = O.loss_and_grads(sd, video, audio, hm)
        finally:
            O.EMULATE_BF16, O.SKIP = False, set()
        num = sum((eg[n] - fg[n]).pow(2).sum().item() for n in fg)
        return {"grad_global_rel": (num / den) ** 0.5, "logits_mean_abs": (elog - flog).abs().mean().item()}

    print("all roundings           ", run([]))
    print("forward roundings only  ", run(["b:" + t for t in TAGS]))
    print("backward roundings only ", run(["f:" + t for t in TAGS]))
    for t in ["w", "xn", "qkv", "pre", "pooled", "o", "h", "wproj", "patchx"]:
        print(f"without forward rounding of {t:7s}", run(["f:" + t]))


if __name__ == "__main__":
    main()
