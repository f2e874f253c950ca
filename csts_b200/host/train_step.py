"""One training step of the CSTS hot path, as the reference's loop performs it
(tools/train_avgaze_net.py:64-109): forward(return_embed) -> [NCE all-gather] -> frame_softmax ->
sim_matrix -> KLDiv + LOSS_ALPHA*EgoNCE -> backward (DDP all-reduce overlapped) -> grad-norm clip ->
AdamW.  The logging/metric collectives of lines 112-128 are outside the hot path.
"""
import torch

from . import distributed as du
from . import losses
from .utils import frame_softmax, sim_matrix


def construct_optimizer(model, cfg):
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
    return torch.optim.AdamW(groups, lr=cfg.SOLVER.BASE_LR, eps=1e-08, weight_decay=cfg.SOLVER.WEIGHT_DECAY, fused=True)


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


def train_step(cfg, model, optimizer, inputs, audio_frames, labels_hm, lr=None):
    """Forward, loss, backward, clip, optimizer step.  Returns the (device) loss tensor; no host sync."""
    if lr is not None:
        for group in optimizer.param_groups:
            group["lr"] = lr
    loss, _, _, _ = compute_loss(cfg, model, inputs, audio_frames, labels_hm)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    if cfg.SOLVER.CLIP_GRAD_L2NORM:
        torch.nn.utils.clip_grad_norm_(model.parameters(), cfg.SOLVER.CLIP_GRAD_L2NORM, foreach=True)
    optimizer.step()
    return loss.detach()
