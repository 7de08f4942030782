"""Pose encoding -> cameras, after the sampling loop (reference: util/camera_transform.py:64-105).

pytorch3d is not a dependency of this package: `PerspectiveCameras` below is a plain container with the
attributes the reference's callers read (R, T, focal_length, principal_point, device, len()).
"""
from __future__ import annotations

import torch


class PerspectiveCameras:
    def __init__(self, focal_length, R, T, principal_point=None, device=None):
        self.focal_length, self.R, self.T = focal_length, R, T
        self.principal_point = torch.zeros_like(focal_length) if principal_point is None else principal_point
        self.device = R.device if device is None else device

    def __len__(self):
        return self.R.shape[0]


def quaternion_to_matrix(q: torch.Tensor) -> torch.Tensor:
    """Real-first quaternion -> rotation, scale 2/|q|^2 (pytorch3d semantics; no normalisation pass)."""
    w, x, y, z = q.unbind(-1)
    s2 = 2.0 / (q * q).sum(-1)
    m = torch.stack(
        (
            1 - s2 * (y * y + z * z), s2 * (x * y - z * w), s2 * (x * z + y * w),
            s2 * (x * y + z * w), 1 - s2 * (x * x + z * z), s2 * (y * z - x * w),
            s2 * (x * z - y * w), s2 * (y * z + x * w), 1 - s2 * (x * x + y * y),
        ),
        -1,
    )
    return m.reshape(q.shape[:-1] + (3, 3))


def pose_encoding_to_camera(
    pose_encoding: torch.Tensor,
    pose_encoding_type: str = "absT_quaR_logFL",
    log_focal_length_bias: float = 1.8,
    min_focal_length: float = 0.1,
    max_focal_length: float = 20,
    return_dict: bool = False,
):
    if pose_encoding_type != "absT_quaR_logFL":
        raise ValueError(f"Unknown pose encoding {pose_encoding_type}")
    flat = pose_encoding.reshape(-1, pose_encoding.shape[-1])
    T = flat[:, :3]
    R = quaternion_to_matrix(flat[:, 3:7])
    fl = torch.clamp((flat[:, 7:9] + log_focal_length_bias).exp(), min=min_focal_length, max=max_focal_length)
    if return_dict:
        return {"focal_length": fl, "R": R, "T": T}
    return PerspectiveCameras(focal_length=fl, R=R, T=T, device=R.device)
