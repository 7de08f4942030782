"""Evaluation metrics of the reference (util/metric.py) on top of the native pair kernel.

`camera_to_rel_deg` (:14-48) runs on the GPU (pdb_rel_pose_error: one thread per camera pair).  `calculate_auc_np` (:51-78),
`calculate_auc` (:81-107) and `compute_ARE` (:182-192) reduce a few hundred angles and are host code in the reference as well
(numpy); they are restated here with the same binning rules.  The demo's optional Umeyama alignment
(pytorch3d `corresponding_cameras_alignment`, demo.py:126-128) is third-party and out of scope.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native


def camera_to_rel_deg(pred_cameras, gt_cameras, device, batch_size):
    """Relative rotation / translation-direction errors in degrees for all camera pairs i < j of each of the `batch_size`
    sequences (order: torch.combinations, sequence major).  Returns CUDA tensors, like the reference on a CUDA device."""
    R_pred, T_pred = torch.as_tensor(pred_cameras.R), torch.as_tensor(pred_cameras.T)
    if not R_pred.is_cuda:
        raise _native.NativeError("cameras must live on a CUDA device (posediffusion_b200 has no CPU fallback)")
    dev = R_pred.device
    R_gt = torch.as_tensor(gt_cameras.R, dtype=torch.float32).to(dev)
    T_gt = torch.as_tensor(gt_cameras.T, dtype=torch.float32).to(dev)
    ctx = _native.Context.get(dev)
    return ctx.rel_pose_error(R_pred, T_pred, R_gt, T_gt, int(batch_size))


def calculate_auc_np(r_error, t_error, max_threshold=30):
    """np.histogram of max(r, t) over integer bins [0, 1, ..., max_threshold], normalised, mean of the cumulative sum."""
    r_error, t_error = np.asarray(r_error), np.asarray(t_error)
    max_errors = np.maximum(r_error, t_error)
    histogram, _ = np.histogram(max_errors, bins=np.arange(max_threshold + 1))
    return np.mean(np.cumsum(histogram.astype(float) / float(len(max_errors))))


def calculate_auc(r_error, t_error, max_threshold=30):
    """The torch.histc variant: max_threshold + 1 equal bins over [0, max_threshold] (values outside are dropped, the upper
    edge belongs to the last bin), float32 arithmetic."""
    r = torch.as_tensor(r_error).detach().float().cpu().numpy()
    t = torch.as_tensor(t_error).detach().float().cpu().numpy()
    max_errors = np.maximum(r, t)
    bins = max_threshold + 1
    inside = max_errors[(max_errors >= 0) & (max_errors <= max_threshold)]
    pos = np.floor((inside - np.float32(0)) / np.float32(max_threshold) * np.float32(bins)).astype(np.int64)
    pos[pos == bins] = bins - 1
    histogram = np.bincount(pos, minlength=bins).astype(np.float32)
    normalized = histogram / np.float32(len(max_errors))
    return torch.tensor(np.cumsum(normalized, dtype=np.float32).mean(dtype=np.float32))


def compute_ARE(rotation1, rotation2):
    """Absolute rotation error in degrees per camera, folded to [0, 90] (min(err, |180 - err|))."""
    if isinstance(rotation1, torch.Tensor):
        rotation1 = rotation1.cpu().detach().numpy()
    if isinstance(rotation2, torch.Tensor):
        rotation2 = rotation2.cpu().detach().numpy()
    R_rel = np.einsum("Bij,Bjk ->Bik", rotation1.transpose(0, 2, 1), rotation2)
    t = (np.trace(R_rel, axis1=1, axis2=2) - 1) / 2
    theta = np.arccos(np.clip(t, -1, 1))
    error = theta * 180 / np.pi
    return np.minimum(error, np.abs(180 - error))
