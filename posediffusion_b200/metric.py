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
    """Mean of the cumulative, pair-normalised histogram of max(r, t) over the integer-degree bins [k, k+1), k < max_threshold
    (np.histogram semantics: the last bin is closed on the right, larger errors are not counted)."""
    worst = np.maximum(np.asarray(r_error), np.asarray(t_error))
    counts = np.histogram(worst, bins=np.arange(max_threshold + 1))[0].astype(float)
    return np.mean(np.cumsum(counts / float(worst.shape[0])))


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


def _as_numpy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def compute_ARE(rotation1, rotation2):
    """Absolute rotation error per camera in degrees: angle of R1^T R2, folded into [0, 90] by min(e, |180 - e|)."""
    R1, R2 = _as_numpy(rotation1), _as_numpy(rotation2)
    cos_angle = (np.einsum("bji,bji->b", R1, R2) - 1.0) / 2.0  # trace(R1^T R2) = sum_ij R1[j,i] R2[j,i]
    degrees = np.degrees(np.arccos(np.clip(cos_angle, -1.0, 1.0)))
    return np.minimum(degrees, np.abs(180.0 - degrees))
