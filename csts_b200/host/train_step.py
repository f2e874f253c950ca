"""One training step of the CSTS hot path, as the reference's loop performs it
(tools/train_avgaze_net.py:64-109): forward(return_embed) -> [NCE all-gather] -> frame_softmax ->
sim_matrix -> KLDiv + LOSS_ALPHA*EgoNCE -> backward (DDP all-reduce overlapped) -> grad-norm clip ->
AdamW.  The logging/metric collectives of lines 112-128 are outside the hot path.
"""
import torch

from . import distributed as du
from . import losses
from .utils import frame_softmax, sim_matrix


def construct_optimizer(model, cfg, capturable=False, fused_clip=False):
    """Parameter grouping of slowfast/models/optimizer.py:11-108 for the AdamW case: weight decay on
    matrices / conv kernels / position embeddings, zero weight decay on 1-D parameters and biases
    (SOLVER.ZERO_WD_1D_PARAM) and on model.no_weight_decay()."""
    inner = model.module if hasattr(model, "module") else model
    skip = set(inner.no_weight_decay()) if hasattr(inner, "no_weight_decay") else set()
    decay, no_decay = [], []
    for name, m in inner.named_modules():
        assert not isinstance(m, torch.nn.modules.batchnorm._NormBase), "CSTS has no BatchNorm"
        for p in m.parameters(recurse=False):
            if not p.requires_grad:
                continue
            if name in skip or (cfg.SOLVER.ZERO_WD_1D_PARAM and (p.dim() == 1 or name.endswith(".bias"))):
                no_decay.append(p)
            else:
                decay.append(p)
    assert len(decay) + len(no_decay) == len([p for p in inner.parameters() if p.requires_grad])
    groups = [g for g in ({"params": decay, "weight_decay": cfg.SOLVER.WEIGHT_DECAY},
                          {"params": no_decay, "weight_decay": 0.0}) if g["params"]]
    if cfg.SOLVER.OPTIMIZING_METHOD != "adamw":
        raise NotImplementedError("the CSTS configs train with AdamW")
    lr = cfg.SOLVER.BASE_LR
    if capturable:      # CUDA-graph replay: the learning rate must live on the device
        lr = torch.tensor(float(lr), dtype=torch.float32, device=decay[0].device)
    if fused_clip:      # unscale + clip + AdamW + 16-bit weight refresh in one pass (host/optimizer.py); graph-capturable
        from .optimizer import FusedClipAdamW
        return FusedClipAdamW(groups, lr=lr, weight_decay=cfg.SOLVER.WEIGHT_DECAY, eps=1e-08, weight_cache=getattr(inner, "_wc", None))
    return torch.optim.AdamW(groups, lr=lr, eps=1e-08, weight_decay=cfg.SOLVER.WEIGHT_DECAY, fused=True, capturable=capturable)


def compute_loss(cfg, model, inputs, audio_frames, labels_hm):
    """Lines 70-92 of the reference loop for MODEL.LOSS_FUNC == 'kldiv+egonce' (and plain 'kldiv')."""
    if cfg.MODEL.LOSS_FUNC == "kldiv+egonce":
        preds, v_embed, a_embed = model(inputs, audio_frames, return_embed=True)
        if du.get_world_size() > 1:
            v_embed, a_embed = du.all_gather_with_grad([v_embed, a_embed])
        preds = frame_softmax(preds, temperature=2)
        similarity = sim_matrix(v_embed, a_embed)
        kldiv_loss = losses.get_loss_func("kldiv")()(preds, labels_hm)
        egonce_loss = losses.get_loss_func("egonce")()(similarity)
        loss = kldiv_loss + cfg.MODEL.LOSS_ALPHA * egonce_loss
        return loss, preds, kldiv_loss, egonce_loss
    if cfg.MODEL.LOSS_FUNC == "kldiv":
        preds = frame_softmax(model(inputs, audio_frames), temperature=2)
        loss = losses.get_loss_func("kldiv")()(preds, labels_hm)
        return loss, preds, loss, None
    raise NotImplementedError(f"loss {cfg.MODEL.LOSS_FUNC} is outside the CSTS hot path")


def make_grad_scaler(cfg, **kw):
    """tools/train_avgaze_net.py:277 — enabled only under TRAIN.MIXED_PRECISION (fp16 storage mode)."""
    return torch.amp.GradScaler("cuda", enabled=bool(cfg.TRAIN.MIXED_PRECISION), **kw)


def check_nan_losses(loss):
    """slowfast/utils/misc.py:26-33 — raises on a NaN loss.  Reads the loss on the host (one sync), which is why the
    hot step only does it on request."""
    import math
    from datetime import datetime
    if math.isnan(float(loss)):
        raise RuntimeError("ERROR: Got NaN losses {}".format(datetime.now()))


def train_step(cfg, model, optimizer, inputs, audio_frames, labels_hm, lr=None, grad_sync=None, scaler=None, check_nan=False):
    """Forward, loss, backward, [gradient exchange], clip, optimizer step.  Returns the (device) loss tensor;
    no host sync unless check_nan (the reference loop's check_nan_losses, tools/train_avgaze_net.py:95; not
    available while capturing a CUDA graph).  With a DDP-wrapped model the exchange happens inside backward (DDP reducer); for an
    un-wrapped replica pass grad_sync (e.g. distributed.allreduce_gradients) to average gradients here.
    `scaler` is the reference loop's GradScaler (lines 99-109: scale(loss).backward(), unscale_, clip, step,
    update); with the fused AdamW none of its calls synchronises the host, so the step stays graph-capturable."""
    if lr is not None:
        for group in optimizer.param_groups:
            if torch.is_tensor(group["lr"]):
                group["lr"].fill_(lr)
            else:
                group["lr"] = lr
    loss, _, _, _ = compute_loss(cfg, model, inputs, audio_frames, labels_hm)
    if check_nan:
        check_nan_losses(loss)
    optimizer.zero_grad(set_to_none=True)
    scaling = scaler is not None and scaler.is_enabled()
    assert scaling or not cfg.TRAIN.MIXED_PRECISION, "TRAIN.MIXED_PRECISION stores fp16 gradients: pass make_grad_scaler(cfg)"
    root = scaler.scale(loss) if scaling else loss
    # This function owns the backward pass of an un-wrapped replica, so the weight-gradient branch (block.py::_Fork) may
    # run un-joined across blocks and is joined once below; under DDP every block joins (the reducer reads gradients
    # from its own hooks).
    wc = getattr(model, "_wc", None)
    if wc is not None:
        wc.defer_join = True
    try:
        if hasattr(grad_sync, "start"):          # OverlappedGradSync: exchange runs during backward
            grad_sync.start()
            root.backward()
            if wc is not None:
                wc.join_backward()
            grad_sync.finish()
        else:
            root.backward()
            if wc is not None:
                wc.join_backward()
            if grad_sync is not None:
                grad_sync()
    finally:
        if wc is not None:
            wc.defer_join = False
    clip_val = getattr(cfg.SOLVER, "CLIP_GRAD_VAL", None)
    if hasattr(optimizer, "clip_and_step"):          # FusedClipAdamW: lines 101-109 of the reference loop in two launches
        if clip_val:
            raise NotImplementedError("SOLVER.CLIP_GRAD_VAL (element-wise clipping, tools/train_avgaze_net.py:103-104) is not "
                                      "fused: construct the optimizer with fused_clip=False")
        optimizer.clip_and_step(cfg.SOLVER.CLIP_GRAD_L2NORM or 0.0, scaler if scaling else None)
        return loss.detach()
    if scaling:
        scaler.unscale_(optimizer)
    if clip_val:                                       # takes priority over the norm clip, as in the reference (:103-106)
        torch.nn.utils.clip_grad_value_(model.parameters(), clip_val)
    elif cfg.SOLVER.CLIP_GRAD_L2NORM:
        torch.nn.utils.clip_grad_norm_(model.parameters(), cfg.SOLVER.CLIP_GRAD_L2NORM, foreach=True)
    if scaling:
        scaler.step(optimizer)
        scaler.update()
    else:
        optimizer.step()
    return loss.detach()


class GraphedTrainStep:
    """The whole training step (weight casts, forward, loss, backward, clip, AdamW) captured once into
    a CUDA graph and replayed: ~1100 kernel launches per step cost one graph launch on the host.

    All shapes of the path are static (fixed clip geometry, fixed batch), which is what makes the
    capture legal.  Inputs are copied into static device buffers before each replay, so the caller
    keeps passing ordinary (host or device) tensors.  The optimizer must be constructed with
    ``construct_optimizer(model, cfg, capturable=True)``.
    """

    def __init__(self, cfg, model, optimizer, video, audio, labels_hm, warmup=3, scaler=None):
        # Data parallel: the replica is stepped un-wrapped and gradients are averaged by captured NCCL
        # all-reduces (DDP's reducer is host-driven and cannot be replayed from a graph); DDP's
        # constructor has already broadcast rank 0's parameters.
        inner = model.module if hasattr(model, "module") else model
        model = inner
        self.grad_sync = None
        if du.get_world_size() > 1:
            import os
            if os.environ.get("CSTS_OVERLAP_ALLREDUCE", "1") == "1":
                self.grad_sync = du.OverlappedGradSync(inner)
            else:
                self.grad_sync = lambda: du.allreduce_gradients(list(inner.parameters()))
        dev = next(inner.parameters()).device
        self.cfg, self.model, self.optimizer = cfg, model, optimizer
        self.scaler = scaler if scaler is not None else make_grad_scaler(cfg)
        self._stage = None
        self.video = torch.empty(video.shape, dtype=torch.float32, device=dev)
        self.audio = torch.empty(audio.shape, dtype=torch.float32, device=dev)
        self.labels = torch.empty(labels_hm.shape, dtype=torch.float32, device=dev)
        self._load(video, audio, labels_hm)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):              # eager warm-up: lazy inits (func attributes, optimizer state, NCCL)
                train_step(cfg, model, optimizer, [self.video], self.audio, self.labels, grad_sync=self.grad_sync, scaler=self.scaler)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # The 16-bit weight copies: with torch's optimizer they are rebuilt at the top of every training forward
        # (WeightCache.begin_training_step), so the casts become part of the captured step; with the fused
        # optimizer the AdamW kernel itself rewrites them at the end of every step and no cast is captured.
        optimizer.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = train_step(cfg, model, optimizer, [self.video], self.audio, self.labels, grad_sync=self.grad_sync,
                                   scaler=self.scaler)

    def _load(self, video, audio, labels_hm):
        self.video.copy_(video[0] if isinstance(video, (list, tuple)) else video, non_blocking=True)
        self.audio.copy_(audio, non_blocking=True)
        self.labels.copy_(labels_hm, non_blocking=True)

    # ---- input pipelining: the next batch travels host -> device while the current step computes ----------------
    def prefetch(self, inputs, audio_frames, labels_hm):
        """Start the asynchronous host->device copy of the NEXT batch (pinned host tensors) into staging buffers on a
        copy stream; `step_prefetched()` consumes it.  What a data loader with a prefetch queue does."""
        if self._stage is None:
            self._stage = tuple(torch.empty_like(t) for t in (self.video, self.audio, self.labels))
            self._copy_stream = torch.cuda.Stream(device=self.video.device)
            self._ev_copied, self._ev_free = torch.cuda.Event(), torch.cuda.Event()
        cs = self._copy_stream
        cs.wait_event(self._ev_free)                 # the previous staged batch has been moved into the step's buffers
        with torch.cuda.stream(cs):
            self._stage[0].copy_(inputs[0] if isinstance(inputs, (list, tuple)) else inputs, non_blocking=True)
            self._stage[1].copy_(audio_frames, non_blocking=True)
            self._stage[2].copy_(labels_hm, non_blocking=True)
            self._ev_copied.record(cs)

    def step_prefetched(self, lr=None):
        """One step on the batch handed to the last prefetch()."""
        assert self._stage is not None, "call prefetch() first"
        cur = torch.cuda.current_stream(self.video.device)
        cur.wait_event(self._ev_copied)
        self.video.copy_(self._stage[0], non_blocking=True)          # device -> device, ~0.03 ms
        self.audio.copy_(self._stage[1], non_blocking=True)
        self.labels.copy_(self._stage[2], non_blocking=True)
        self._ev_free.record(cur)
        return self(None, None, None, lr=lr)

    def __call__(self, inputs, audio_frames, labels_hm, lr=None):
        if lr is not None:
            for group in self.optimizer.param_groups:
                group["lr"].fill_(lr)
        if inputs is not None:
            self._load(inputs, audio_frames, labels_hm)
        self.graph.replay()
        if not hasattr(self.optimizer, "clip_and_step"):
            # torch's optimizer moved the parameters; the captured casts refresh the 16-bit copies only at the top of the
            # NEXT replay, so an eager (evaluation) forward in between must rebuild them
            self.model._wc.invalidate()
        return self.loss
