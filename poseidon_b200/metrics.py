"""Evaluation metrics computed on the device.

Same functions and semantics as the reference's `scOT/metrics.py` (:4-56) and the `compute_metrics` closure of
`scOT/train.py:344-398`, but the per-pixel reductions run in one kernel of libscot_b200.so on the GPU-resident
predictions: only 2 floats per (sample, channel) cross PCIe instead of the whole prediction array. Inputs are CUDA
tensors [N, C, H, W]; results are numpy arrays / floats like the reference's. No CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib


def _plane_sums(preds: torch.Tensor, targets: torch.Tensor, p: int) -> np.ndarray:
    if not (preds.is_cuda and targets.is_cuda):
        raise RuntimeError("poseidon_b200.metrics works on CUDA tensors only (no CPU fallback)")
    if preds.shape != targets.shape or preds.dim() != 4:
        raise ValueError("preds / targets must both be [N, C, H, W]")
    a = preds.detach().to(torch.float32).contiguous()
    b = targets.detach().to(torch.float32).contiguous()
    n, c, h, w = a.shape
    out = torch.empty(n * c * 2, device=a.device, dtype=torch.float32)
    _lib.check(_lib.load().scot_lp_plane_sums(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), int(p), n * c, h * w, _lib.cur_stream()),
               "scot_lp_plane_sums")
    return out.cpu().numpy().astype(np.float64).reshape(n, c, 2)


def lp_error(preds, targets, p=1):
    """scOT/metrics.py:4-9"""
    s = _plane_sums(preds, targets, p)
    return np.sum(s[:, :, 0], axis=-1) ** (1 / p)


def relative_lp_error(preds, targets, p=1, return_percent=True):
    """scOT/metrics.py:12-36: per sample, (sum_c,px |pred-y|^p / sum_c,px |y|^p)^(1/p) [* 100]"""
    s = _plane_sums(preds, targets, p)
    norm = np.sum(s[:, :, 1], axis=-1)
    norm = np.where(norm == 0, 1e-10, norm)
    errors = (np.sum(s[:, :, 0], axis=-1) / norm) ** (1 / p)
    if return_percent:
        errors = errors * 100
    return errors


def mean_relative_lp_error(preds, targets, p=1, return_percent=True):
    return np.mean(relative_lp_error(preds, targets, p, return_percent), axis=0)


def median_relative_lp_error(preds, targets, p=1, return_percent=True):
    return np.median(relative_lp_error(preds, targets, p, return_percent), axis=0)


def error_statistics(preds: torch.Tensor, targets: torch.Tensor, channel_slice_list: Sequence[int],
                     printable_channel_description: Optional[List[str]] = None) -> Dict[str, float]:
    """The `compute_metrics` closure of scOT/train.py:344-398 for device-resident predictions: relative L1 error (in
    percent) per channel group -> median / mean / std / min / max over the samples, plus the means over the groups."""
    s = _plane_sums(preds, targets, 1)
    groups = []
    for i in range(len(channel_slice_list) - 1):
        lo, hi = channel_slice_list[i], channel_slice_list[i + 1]
        norm = np.sum(s[:, lo:hi, 1], axis=-1)
        norm = np.where(norm == 0, 1e-10, norm)
        e = np.sum(s[:, lo:hi, 0], axis=-1) / norm * 100
        groups.append({"median_relative_l1_error": np.median(e, axis=0), "mean_relative_l1_error": np.mean(e, axis=0),
                       "std_relative_l1_error": np.std(e, axis=0), "min_relative_l1_error": np.min(e, axis=0),
                       "max_relative_l1_error": np.max(e, axis=0)})
    if len(groups) == 1:
        return groups[0]
    names = printable_channel_description or [f"group{i}" for i in range(len(groups))]
    out = {"mean_relative_l1_error": np.mean(np.array([g["mean_relative_l1_error"] for g in groups]), axis=0),
           "mean_over_median_relative_l1_error": np.mean(np.array([g["median_relative_l1_error"] for g in groups]), axis=0)}
    for name, g in zip(names, groups):
        for k, v in g.items():
            out[name + "/" + k] = v
    return out
