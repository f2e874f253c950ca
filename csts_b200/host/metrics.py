"""Adaptive-threshold F1 of the gaze heat-maps on the GPU.

Mirror of ``slowfast/utils/metrics.py:9-74`` — same arguments, same threshold grids per dataset, same return value —
backed by one fused kernel pair (csts_adaptive_f1) instead of the reference's two (thresholds, B, T, H, W) f32 temporaries
and ~100 element-wise launches.  ``adaptive_f1_async`` returns the device tensor (f1, recall, precision, threshold, index)
without a host round trip, and can fuse the loops' per-frame min-max rescale (tools/train_avgaze_net.py:125-127,
tools/test_avgaze_net.py:66-68), so a training loop can evaluate the metric every LOG_PERIOD iterations and read it back
together with the losses.
"""
import numpy as np
import torch

from .. import _lib

_FIX0 = ("ego4dgaze", "ego4dgaze_forecast", "ego4d_av_gaze", "ego4d_av_gaze_forecast", "aria_gaze", "aria_gaze_forecast", "aria_av_gaze",
         "aria_av_gaze_forecast")


def thresholds_for(dataset):
    """metrics.py:31-39 — the search grid depends on the dataset."""
    if "forecast" in dataset and "aria" not in dataset:
        return np.linspace(0.01, 0.07, 31)
    if "forecast" in dataset and "aria" in dataset:
        return np.linspace(0.0, 0.02, 21)
    return np.linspace(0, 0.02, 11)


def fixation_index(dataset):
    """metrics.py:52-58."""
    if dataset == "egteagaze":
        return 1
    if dataset in _FIX0:
        return 0
    raise NotImplementedError(f"Metrics of {dataset} is not implemented.")


_thr_cache = {}


def adaptive_f1_async(preds, labels_hm, labels, dataset, rescale=False):
    """preds (B, 1, T, H, W) or (B, T, H, W) f32, labels_hm (B, T, H, W), labels (B, T, 3) on the GPU.
    Returns a device tensor [f1, recall, precision, threshold, threshold index]; nothing is synchronised.
    rescale=True applies (p - min) / (max - min + 1e-6) per frame first (what the loops do before calling the metric)."""
    thr = thresholds_for(dataset)
    key = (dataset, preds.device)
    if key not in _thr_cache:
        # `preds > thresholds[i]` compares in the tensor's dtype: the float64 grid value is rounded to f32
        _thr_cache[key] = torch.tensor(thr.astype(np.float32), device=preds.device)
    thr_d = _thr_cache[key]
    p = preds.detach().contiguous().float()
    hm = labels_hm.contiguous().float()
    lb = labels.contiguous().float()
    HW = p.shape[-1] * p.shape[-2]
    frames = p.numel() // HW
    assert hm.numel() == p.numel() and lb.numel() == 3 * frames, "adaptive_f1: preds / labels_hm / labels disagree on (B, T)"
    counts = torch.empty(frames * (2 * len(thr) + 1), dtype=torch.float32, device=p.device)
    out = torch.empty(5, dtype=torch.float32, device=p.device)
    _lib.call("csts_adaptive_f1", _lib.ptr(p), _lib.ptr(hm), _lib.ptr(lb), _lib.ptr(thr_d), len(thr), frames, HW, fixation_index(dataset),
              int(bool(rescale)), _lib.ptr(counts), _lib.ptr(out))
    return out


def adaptive_f1(preds, labels_hm, labels, dataset):
    """Drop-in for slowfast.utils.metrics.adaptive_f1: (f1, recall, precision, threshold) as Python floats
    (the threshold is the float64 grid value, as in the reference).  One host read."""
    out = adaptive_f1_async(preds, labels_hm, labels, dataset).cpu()
    return float(out[0]), float(out[1]), float(out[2]), thresholds_for(dataset)[int(out[4])]
