"""BASELINE.json configs[3]: CSTS_Aria_Gaze_Forecast.yaml inference (model.eval(), no_grad) batch sweep on
one B200.  Forward clips/s per batch size, CUDA-event timed, L2 flushed between iterations."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import csts_oracle as O  # noqa: E402
from csts_b200.host.build import build_model  # noqa: E402
from csts_b200.host.config import get_cfg  # noqa: E402
from csts_b200.host.utils import frame_softmax  # noqa: E402


def main():
    batches = [int(b) for b in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64, 128]
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(ROOT, "configs", "Aria", "CSTS_Aria_Gaze_Forecast.yaml"))
    cfg.merge_from_list(["NUM_GPUS", 1])
    torch.manual_seed(cfg.RNG_SEED)
    model = build_model(cfg).eval()
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    for B in batches:
        video, audio, _ = O.synthetic_batch(min(B, 8), seed=3)
        reps = (B + 7) // 8
        video = video.repeat(reps, 1, 1, 1, 1)[:B].cuda()
        audio = audio.repeat(reps, 1, 1, 1, 1)[:B].cuda()
        with torch.no_grad():
            def step():
                return frame_softmax(model([video], audio), temperature=2)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()            # ~450 launches per forward: replay as one graph
            with torch.cuda.graph(g):
                out = step()
            ts = []
            for _ in range(5):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                g.replay()
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
        ts.sort()
        ms = ts[len(ts) // 2]
        print(json.dumps({"config": "CSTS_Aria_Gaze_Forecast inference", "batch": B, "ms": round(ms, 3), "clips_per_s": round(B / ms * 1e3, 1),
                          "fwd_tflops": round(B * 208.09e9 / (ms * 1e-3) / 1e12, 1), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}),
              flush=True)
        del g, out


if __name__ == "__main__":
    main()
