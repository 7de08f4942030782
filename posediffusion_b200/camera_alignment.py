"""`corresponding_cameras_alignment` as the reference's demo calls it before the absolute rotation error
(pose_diffusion/demo.py:24,126-128 -> pytorch3d.ops.corresponding_cameras_alignment; pose_diffusion/test.py:21 imports it too).

pytorch3d is not a dependency of this package: the published algorithm of its "extrinsics" mode is restated in
csrc/align.cuh and runs in the native library (pdb_cameras_align: one warp for the estimate, one thread per camera for
the application).  There is no CPU path.
"""
from __future__ import annotations

import torch

from . import _native
from .camera_transform import PerspectiveCameras


def corresponding_cameras_alignment(cameras_src, cameras_tgt, estimate_scale: bool = True, mode: str = "extrinsics", eps: float = 1e-9):
    """Aligned copy of `cameras_src` (R, T replaced; focal length / principal point kept), like pytorch3d's function."""
    R_src, T_src = torch.as_tensor(cameras_src.R), torch.as_tensor(cameras_src.T)
    if tuple(R_src.shape) != tuple(torch.as_tensor(cameras_tgt.R).shape):
        raise ValueError("cameras_src and cameras_tgt have to contain the same number of cameras!")
    if mode == "centers":
        raise NotImplementedError("mode='centers' is not on the reference's path (demo.py:128 uses 'extrinsics')")
    if mode != "extrinsics":
        raise ValueError("mode has to be one of (centers, extrinsics)")
    if not R_src.is_cuda:
        raise _native.NativeError("cameras must live on a CUDA device (posediffusion_b200 has no CPU fallback)")
    dev = R_src.device
    R_tgt = torch.as_tensor(cameras_tgt.R, dtype=torch.float32).to(dev)
    T_tgt = torch.as_tensor(cameras_tgt.T, dtype=torch.float32).to(dev)
    R, T, _ = _native.Context.get(dev).cameras_align(R_src, T_src, R_tgt, T_tgt, estimate_scale, eps)
    return PerspectiveCameras(focal_length=cameras_src.focal_length, R=R, T=T,
                              principal_point=getattr(cameras_src, "principal_point", None), device=dev)
