#!/usr/bin/env python3
"""Gradient / loss deviation from the fp32 reference, side by side (VERDICT r1 "next" 3-i):

  * the reference itself under torch.autocast(bfloat16) and torch.autocast(float16) — stock PyTorch mixed precision,
  * csts_b200 in its bf16 mode and in its fp16 (TRAIN.MIXED_PRECISION) mode,

all against the reference in fp32 (TF32 off) on the same weights and the same synthetic batch, on one GPU.
Writes gpurun_out/parity_vs_autocast.json.   python tools/parity_report.py [--batch 2]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    import torch
    import csts_oracle as O
    import ref_train
    from csts_b200.host.build import build_model
    from csts_b200.host.config import get_cfg
    from csts_b200.host.train_step import compute_loss
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", 0)
    with open(os.path.join(ROOT, "tests", "golden", "param_shapes.json")) as f:
        shapes = json.load(f)
    sd = O.synthetic_state(shapes, seed=args.seed)
    video, audio, hm = (t.to(dev) for t in O.synthetic_batch(args.batch, seed=args.seed + 1))

    def ref_run(dtype):
        st = ref_train.ReferenceStepper(dev, autocast_dtype=dtype, droppath=0.0, state_dict=sd)
        if st.scaler.is_enabled():
            st.scaler = torch.amp.GradScaler("cuda", init_scale=4096.0)
        loss = st.step(video, audio, hm, optimize=False)
        scale = st.scaler.get_scale() if st.scaler.is_enabled() else 1.0
        grads = {n: p.grad.detach().float() / scale for n, p in st.named_parameters().items()}
        return loss.item(), grads, st.kind

    def ours_run(mixed):
        cfg = get_cfg()
        cfg.merge_from_file(os.path.join(ROOT, "configs", "Ego4D", "CSTS_Ego4D_Gaze_Forecast.yaml"))
        cfg.merge_from_list(["NUM_GPUS", 1, "MODEL.LOSS_FUNC", "kldiv+egonce", "MVIT.DROPPATH_RATE", 0.0, "TRAIN.MIXED_PRECISION", mixed])
        m = build_model(cfg)
        m.load_state_dict(sd, strict=True)
        m.train()
        loss, _, _, _ = compute_loss(cfg, m, [video], audio, hm)
        scale = 4096.0 if mixed else 1.0
        (loss * scale).backward()
        return loss.item(), {n: p.grad.detach().float() / scale for n, p in m.named_parameters()}

    f_loss, f_grads, kind = ref_run(None)

    def compare(loss, grads):
        num = sum((grads[n] - f_grads[n]).pow(2).sum().item() for n in f_grads)
        den = sum(g.pow(2).sum().item() for g in f_grads.values())
        per = sorted(((grads[n] - g).norm() / g.norm()).item() for n, g in f_grads.items() if g.norm() > 1e-6)
        return {"loss_rel": abs(loss - f_loss) / abs(f_loss), "grad_global_rel": (num / den) ** 0.5, "grad_median": per[len(per) // 2],
                "grad_worst": per[-1], "tensors_over_2e-2": sum(1 for e in per if e > 2e-2), "tensors": len(per)}

    report = {"batch": args.batch, "reference_kind": kind, "fp32_loss": f_loss,
              "what": "deviation from the fp32 reference (TF32 off), same weights / inputs, DROPPATH 0; BASELINE tolerance: loss 1e-3, gradients 2e-2"}
    report["reference_autocast_bf16"] = compare(*ref_run(torch.bfloat16)[:2])
    report["reference_autocast_fp16"] = compare(*ref_run(torch.float16)[:2])
    report["csts_b200_bf16"] = compare(*ours_run(False))
    report["csts_b200_fp16"] = compare(*ours_run(True))
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_vs_autocast.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
