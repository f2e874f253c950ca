#!/usr/bin/env python3
"""Run-to-run reproducibility of one forward + backward (same model, same batch, twice): loss / logits / per-tensor
gradient differences.  Diagnostic for stream races and for the amplification of f32 summation-order noise.
    python tools/grad_noise.py [--batch 4] [--mixed]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--mixed", action="store_true")
    args = ap.parse_args()
    import csts_oracle as O
    from csts_b200.host.build import build_model
    from csts_b200.host.config import get_cfg
    from csts_b200.host.train_step import compute_loss
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(ROOT, "configs", "Ego4D", "CSTS_Ego4D_Gaze_Forecast.yaml"))
    cfg.merge_from_list(["NUM_GPUS", 1, "MODEL.LOSS_FUNC", "kldiv+egonce", "MVIT.DROPPATH_RATE", 0.0, "TRAIN.MIXED_PRECISION", args.mixed])
    shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "param_shapes.json")))
    sd = O.synthetic_state(shapes, seed=0)
    m = build_model(cfg)
    m.load_state_dict(sd)
    m.train()
    v, a, h = (t.cuda() for t in O.synthetic_batch(args.batch, seed=1))
    scale = 4096.0 if args.mixed else 1.0

    def run():
        for p in m.parameters():
            p.grad = None
        loss, preds, _, _ = compute_loss(cfg, m, [v], a, h)
        (loss * scale).backward()
        torch.cuda.synchronize()
        return loss.item(), preds.detach().clone(), {n: p.grad.clone() for n, p in m.named_parameters()}

    l1, p1, g1 = run()
    l2, p2, g2 = run()
    num = sum((g1[n] - g2[n]).pow(2).sum().item() for n in g1)
    den = sum(g2[n].pow(2).sum().item() for n in g1)
    per = sorted(((((g1[n] - g2[n]).norm() / g2[n].norm().clamp_min(1e-30)).item(), n, g2[n].norm().item()) for n in g1), reverse=True)
    print(json.dumps({"env": {k: os.environ.get(k) for k in ("CSTS_FORK_WGRAD", "CSTS_PARALLEL_AUDIO", "CSTS_GRAD_ARENA")},
                      "loss": [l1, l2], "preds_max_abs_diff": (p1 - p2).abs().max().item(), "grad_global_rel": (num / den) ** 0.5,
                      "bitwise_equal_tensors": sum(1 for n in g1 if torch.equal(g1[n], g2[n])), "tensors": len(g1),
                      "worst": [(round(e, 4), n, float("%.3g" % nn)) for e, n, nn in per[:6]]}))


if __name__ == "__main__":
    main()
