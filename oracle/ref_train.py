"""The reference's training step, driven on CPU or GPU — TEST / BENCHMARK INFRASTRUCTURE ONLY.

``ReferenceStepper`` runs the literal step of ``tools/train_avgaze_net.py:64-109`` (autocast, forward with
return_embed, frame_softmax, sim_matrix, KLDiv + LOSS_ALPHA*EgoNCE, zero_grad, backward, unscale, clip_grad_norm_,
optimizer step) on

  kind "reference": the UNMODIFIED reference — ``slowfast.models.build_model(cfg)`` and
                    ``slowfast.models.optimizer.construct_optimizer`` imported through oracle/ref_shim.py from
                    /root/reference or from baseline/_ref (the pip-installed package, which travels to the GPU box);
  kind "port":      oracle/csts_oracle.py (the functional fp32 restatement) with torch.optim.AdamW in the
                    reference's parameter grouping — used only when no reference tree can be imported.

It is what ``bench.py --impl reference`` (host cores), the ``cpu_baseline`` leg and the ``gpu_reference`` leg
(same B200, stock PyTorch eager: the bar SURVEY.md §8(d) names) time, and what the binding tests compare
against.  Nothing under ``csts_b200/`` imports this module.
"""
import json
import os

import torch

import csts_oracle as O
import ref_shim

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _egonce_on_device(sim, temperature=0.05):
    """EgoNCE.forward (losses.py:157-170) with the eye mask on sim.device — the reference hard-codes .cuda()
    (:158), which cannot run on a CPU-only arm."""
    import torch.nn.functional as F
    mask = torch.eye(sim.shape[0], device=sim.device) > 0
    i_sm = F.softmax(sim / temperature, dim=1)
    j_sm = F.softmax(sim.t() / temperature, dim=1)
    idiag = torch.log(torch.sum(i_sm * mask, dim=1))
    jdiag = torch.log(torch.sum(j_sm * mask, dim=1))
    return -idiag.sum() / len(idiag) - jdiag.sum() / len(jdiag)


class ReferenceStepper:
    def __init__(self, device="cpu", autocast_dtype=None, droppath=0.2, seed=0, state_dict=None, lr=None, force_port=False):
        self.device = torch.device(device)
        self.autocast_dtype = autocast_dtype
        cuda = self.device.type == "cuda"
        self.kind = "reference" if ref_shim.reference_available() and not force_port else "port"
        if self.kind == "reference":
            over = ["NUM_GPUS", 1 if cuda else 0, "MVIT.DROPPATH_RATE", droppath]
            if lr is not None:
                over += ["SOLVER.BASE_LR", lr]
            self.cfg = ref_shim.reference_cfg(overrides=over)
            from slowfast.models import build_model, losses
            from slowfast.models import optimizer as optim
            from slowfast.utils.utils import frame_softmax, sim_matrix
            torch.manual_seed(seed)
            self.model = build_model(self.cfg)            # .cuda() happens inside when NUM_GPUS > 0 (build.py:36-41)
            if state_dict is not None:
                self.model.load_state_dict(state_dict, strict=True)
            self.model.train()
            self.optimizer = optim.construct_optimizer(self.model, self.cfg)
            self._frame_softmax, self._sim_matrix = frame_softmax, sim_matrix
            self._kldiv = losses.get_loss_func("kldiv")().to(self.device)
            self._egonce = losses.get_loss_func("egonce")() if cuda else _egonce_on_device
            self.alpha = self.cfg.MODEL.LOSS_ALPHA
            self.clip = self.cfg.SOLVER.CLIP_GRAD_L2NORM
        else:
            with open(os.path.join(_REPO, "tests", "golden", "param_shapes.json")) as f:
                shapes = json.load(f)
            sd = state_dict if state_dict is not None else O.synthetic_state(shapes, seed=seed)
            self.params = {k: torch.nn.Parameter(v.detach().clone().to(self.device)) for k, v in sd.items()}
            decay = [p for p in self.params.values() if p.dim() > 1]
            no_decay = [p for p in self.params.values() if p.dim() <= 1]
            self.optimizer = torch.optim.AdamW([{"params": decay, "weight_decay": 0.05}, {"params": no_decay, "weight_decay": 0.0}],
                                               lr=lr if lr is not None else 1e-4, eps=1e-8)
            self.alpha, self.clip = 0.05, 1.0
        self.scaler = torch.amp.GradScaler(self.device.type, enabled=autocast_dtype == torch.float16)

    def parameters(self):
        return list(self.model.parameters()) if self.kind == "reference" else list(self.params.values())

    def named_parameters(self):
        return dict(self.model.named_parameters()) if self.kind == "reference" else dict(self.params)

    def loss(self, video, audio, labels_hm):
        """Lines 70-88."""
        with torch.autocast(self.device.type, dtype=self.autocast_dtype or torch.bfloat16, enabled=self.autocast_dtype is not None):
            if self.kind == "reference":
                preds, v, a = self.model([video], audio, return_embed=True)
                preds = self._frame_softmax(preds, temperature=2)
                sim = self._sim_matrix(v, a)
                kld = self._kldiv(preds, labels_hm)
                nce = self._egonce(sim)
            else:
                logits, v, a = O.csts_forward(self.params, video, audio, return_embed=True)
                preds = O.frame_softmax(logits, 2.0)
                kld = O.kldiv(preds, labels_hm)
                nce = O.egonce(O.sim_matrix(v, a))
            return kld + self.alpha * nce, kld, nce, preds

    def step(self, video, audio, labels_hm, optimize=True):
        """Lines 70-109.  Returns the (device) loss tensor; no host sync."""
        loss, _, _, _ = self.loss(video, audio, labels_hm)
        self.optimizer.zero_grad()
        self.scaler.scale(loss).backward()
        if not optimize:
            return loss.detach()
        self.scaler.unscale_(self.optimizer)
        if self.clip:
            torch.nn.utils.clip_grad_norm_(self.parameters(), self.clip)
        self.scaler.step(self.optimizer)
        self.scaler.update()
        return loss.detach()


def time_steps(stepper, batch, steps, warmup, seed=100):
    """(clips/s, [seconds per step]) for `steps` timed training steps at `batch` clips, after `warmup` untimed ones.
    CUDA: events on the current stream; CPU: perf_counter."""
    import statistics
    import time
    video, audio, hm = (t.to(stepper.device) for t in O.synthetic_batch(batch, seed=seed))
    cuda = stepper.device.type == "cuda"
    times = []
    for i in range(warmup + steps):
        if cuda:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            stepper.step(video, audio, hm)
            e.record()
            torch.cuda.synchronize()
            dt = s.elapsed_time(e) * 1e-3
        else:
            t0 = time.perf_counter()
            stepper.step(video, audio, hm)
            dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return batch / statistics.median(times), times
