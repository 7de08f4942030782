"""TEST INFRASTRUCTURE (never imported by the product): CPU restatement of pytorch3d.ops.corresponding_cameras_alignment,
the third-party "Umeyama" step of the reference's demo (pose_diffusion/demo.py:126-128, mode="extrinsics",
estimate_scale=True, eps=1e-9).

PARITY UNPINNED: pytorch3d is absent from /root/reference and from this image and its version is not pinned by the
reference (install.sh:24), so there is no reference output to pin this file against.  It restates the published algorithm
(pytorch3d/ops/cameras_alignment.py: `_align_camera_extrinsics` + the application step) with torch's own SVD, and the tests
anchor it on the property that defines the operation: cameras that differ from the targets by a world similarity
transform are mapped back onto the targets exactly.
"""
from __future__ import annotations

import torch


def align_camera_extrinsics(R_src, T_src, R_tgt, T_tgt, estimate_scale: bool = True, eps: float = 1e-9):
    """-> (align_R [3,3], align_T [3], scale) in pytorch3d's row-vector convention X_view = X_world R + T."""
    RRcov = torch.bmm(R_src, R_tgt.transpose(2, 1)).mean(0)
    U, _, Vh = torch.linalg.svd(RRcov)
    align_R = Vh.transpose(0, 1) @ U.transpose(0, 1)  # V @ U.t()
    A = torch.bmm(R_src, T_src[:, :, None])[:, :, 0]
    B = torch.bmm(R_src, T_tgt[:, :, None])[:, :, 0]
    Amu, Bmu = A.mean(0, keepdim=True), B.mean(0, keepdim=True)
    if estimate_scale and A.shape[0] > 1:
        Ac, Bc = A - Amu, B - Bmu
        scale = (Ac * Bc).mean() / (Ac**2).mean().clamp(eps)
    else:
        scale = torch.ones((), dtype=R_src.dtype)
    align_T = (Bmu - scale * Amu)[0]
    return align_R, align_T, scale


def corresponding_cameras_alignment(R_src, T_src, R_tgt, T_tgt, estimate_scale: bool = True, mode: str = "extrinsics", eps: float = 1e-9):
    """-> (R_aligned [N,3,3], T_aligned [N,3]) of the source cameras."""
    if R_src.shape != R_tgt.shape:
        raise ValueError("cameras_src and cameras_tgt have to contain the same number of cameras!")
    if mode != "extrinsics":
        raise ValueError("only mode='extrinsics' (the reference's call) is restated")
    align_R, align_T, scale = align_camera_extrinsics(R_src, T_src, R_tgt, T_tgt, estimate_scale, eps)
    R_new = torch.bmm(align_R[None].expand_as(R_src), R_src)
    T_new = torch.bmm(align_T[None, None].repeat(R_src.shape[0], 1, 1), R_src)[:, 0] + T_src * scale
    return R_new, T_new
