"""Checkpoint load / save for the drop-in model — the part of ``slowfast/utils/checkpoint.py`` a CSTS run needs
(SURVEY.md §8f rank 4): the ``{epoch, model_state, optimizer_state, cfg[, scaler_state]}`` file layout (:110-143),
resume, and the fine-tune load that matches pre-trained weights by NAME AND SHAPE and bilinearly interpolates the
separable position embeddings of a pre-trained MViT (:290-354) — how the reference starts from K400 MViT-B
(224-pixel crops, 16 frames: ``pos_embed_spatial`` (1, 3136, 96), ``pos_embed_temporal`` (1, 8, 96)) for its
256-pixel, 8-frame clips ((1, 4096, 96), (1, 4, 96)).

Caffe2 conversion, 2D->3D inflation and the sub-BatchNorm renames of the reference are not carried over: CSTS has no
BatchNorm and no shipped config uses them (TRAIN.CHECKPOINT_TYPE pytorch, CHECKPOINT_INFLATE False).
"""
import os
from collections import OrderedDict

import torch
import torch.nn.functional as F

INTERPOLATE_PARAMS = ("pos_embed_spatial", "pos_embed_temporal")      # checkpoint.py:327


def match_pretrained_state(pre_train_dict, model_dict, clear_name_pattern=()):
    """checkpoint.py:299-335.  Returns (state to load, names of model tensors left at their initial values)."""
    for item in clear_name_pattern or ():                               # :299-309
        renamed = OrderedDict()
        for k, v in pre_train_dict.items():
            renamed[k.replace(item, "") if item in k else k] = v
        pre_train_dict = renamed
    match = {k: v for k, v in pre_train_dict.items() if k in model_dict and v.size() == model_dict[k].size()}      # :314-318
    not_loaded = [k for k in model_dict.keys() if k not in match]                                                   # :320-324
    for k in INTERPOLATE_PARAMS:                                                                                    # :327-335
        if k in not_loaded and k in model_dict and k in pre_train_dict:
            v = pre_train_dict[k]
            t = model_dict[k].size()
            # (1, L, C) -> (1, 1, L, C): "bilinear" over (length, channel); the channel extent is unchanged, so this is a
            # linear resampling of the position axis with align_corners=False
            match[k] = F.interpolate(v.unsqueeze(0), (t[1], t[2]), mode="bilinear").squeeze(0)
            not_loaded.remove(k)
    return match, not_loaded


def _bare(model, data_parallel):
    return model.module if data_parallel and hasattr(model, "module") else model


def load_checkpoint(path_to_checkpoint, model, data_parallel=True, optimizer=None, scaler=None, epoch_reset=False,
                    clear_name_pattern=(), inflation=False, convert_from_caffe2=False):
    """Signature and return value (the checkpoint's epoch, or -1) of checkpoint.py:185-354."""
    assert os.path.exists(path_to_checkpoint), "Checkpoint '{}' not found".format(path_to_checkpoint)
    if inflation or convert_from_caffe2:
        raise NotImplementedError("caffe2 / inflated checkpoints are outside the CSTS path")
    ms = _bare(model, data_parallel)
    checkpoint = torch.load(path_to_checkpoint, map_location="cpu", weights_only=False)
    match, not_loaded = match_pretrained_state(checkpoint["model_state"], ms.state_dict(), clear_name_pattern)
    ms.load_state_dict(match, strict=False)
    wc = getattr(ms, "_wc", None)
    if wc is not None:
        wc.invalidate()                     # the 16-bit operand copies no longer mirror the parameters
    epoch = -1
    if "epoch" in checkpoint and not epoch_reset:                        # :343-349
        epoch = checkpoint["epoch"]
        if optimizer:
            optimizer.load_state_dict(checkpoint["optimizer_state"])
        if scaler:
            scaler.load_state_dict(checkpoint["scaler_state"])
    load_checkpoint.last_not_loaded = not_loaded
    return epoch


def save_checkpoint(path_to_checkpoint, model, optimizer, epoch, cfg, scaler=None, data_parallel=True):
    """checkpoint.py:110-143 — same record, so the reference's loader reads it and vice versa."""
    ms = _bare(model, data_parallel)
    checkpoint = {"epoch": epoch, "model_state": ms.state_dict(), "optimizer_state": optimizer.state_dict(),
                  "cfg": cfg.dump() if hasattr(cfg, "dump") else cfg}
    if scaler is not None:
        checkpoint["scaler_state"] = scaler.state_dict()
    os.makedirs(os.path.dirname(os.path.abspath(path_to_checkpoint)), exist_ok=True)
    torch.save(checkpoint, path_to_checkpoint)
    return path_to_checkpoint
