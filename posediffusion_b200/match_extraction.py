"""Match ingestion: the reference's COLMAP -> crop-pixel remap (util/match_extraction.py:50-77).

`colmap_keypoint_to_pytorch3d(matches, keypoints, image_info)` keeps the reference's signature and returns the
(kp1, kp2, i12) numpy triple of `matches_dict`; `pack_colmap_matches` feeds the same tables straight to the native packer
(`pdb_matches_pack_colmap`), which fuses the remap into the gather so the 48 B/match arrays are never materialised.
SuperPoint/SuperGlue matching itself (hloc, pycolmap) is upstream of the hot path and not part of this package.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np

from . import _native


def colmap_keypoint_to_pytorch3d(matches: Dict, keypoints: Dict, image_info: Dict):
    """keypoints: {colmap_id (1-based): [k, 2]}, matches: {(r_id, q_id): [m, 2] index pairs or None}."""
    bbox, scale = np.asarray(image_info["bboxes_xyxy"]), np.asarray(image_info["resized_scales"])
    remapped = {}
    for idx, pts in keypoints.items():
        shifted = np.asarray(pts) - 0.5                      # COLMAP pixel centres -> OpenCV
        shifted = shifted - np.asarray([bbox[idx - 1][0], bbox[idx - 1][1]])  # into the centre crop
        remapped[idx] = shifted * scale[idx - 1]             # to the resized (224^2) image
    rows1, rows2, rows12 = [], [], []
    for (r_id, q_id), pair in matches.items():
        if pair is None:
            continue
        rows1.append(remapped[r_id][pair[:, 0]])
        rows2.append(remapped[q_id][pair[:, 1]])
        rows12.append(np.repeat(np.array([[r_id - 1, q_id - 1]]), len(pair), axis=0))
    if not rows1:
        return None, None, None
    return np.concatenate(rows1, 0), np.concatenate(rows2, 0), np.concatenate(rows12, 0)


def pack_colmap_matches(ctx: "_native.Context", matches: Dict, keypoints: Dict, image_info: Dict, img_shape: Tuple[int, int, int, int]):
    """COLMAP tables -> device-resident packed matches (one call, remap fused).  img_shape = (frames, 3, H, W)."""
    frames, _, height, width = (int(v) for v in img_shape)
    n_images = max(keypoints) if keypoints else 0
    is_f64 = all(np.asarray(v).dtype == np.float64 for v in keypoints.values())
    dtype = np.float64 if is_f64 else np.float32
    kp_arrays = [np.ascontiguousarray(keypoints.get(i + 1, np.zeros((0, 2))), dtype=dtype).reshape(-1, 2) for i in range(n_images)]
    kp_ptrs = (C.c_void_p * max(n_images, 1))(*[a.ctypes.data for a in kp_arrays])
    kp_counts = np.asarray([len(a) for a in kp_arrays] or [0], dtype=np.int32)
    pair_ids = np.asarray([[r, q] for (r, q) in matches] or np.zeros((0, 2)), dtype=np.int32).reshape(-1, 2)
    pair_arrays = [None if m is None else np.ascontiguousarray(m, dtype=np.int32).reshape(-1, 2) for m in matches.values()]
    pair_ptrs = (C.c_void_p * max(len(pair_arrays), 1))(*[None if a is None else a.ctypes.data for a in pair_arrays])
    counts = np.asarray([0 if a is None else len(a) for a in pair_arrays] or [0], dtype=np.int32)
    bbox = np.ascontiguousarray(image_info["bboxes_xyxy"], dtype=np.float64).reshape(-1, 4)
    scale = np.ascontiguousarray(image_info["resized_scales"], dtype=np.float64).reshape(-1)
    handle = C.c_void_p()
    fn = ctx.lib.pdb_matches_pack_colmap
    rc = fn(ctx.handle, n_images, kp_ptrs, kp_counts.ctypes.data, int(is_f64), len(pair_arrays), pair_ids.ctypes.data, pair_ptrs,
            counts.ctypes.data, bbox.ctypes.data, scale.ctypes.data, frames, height, width, _native._stream_ptr(ctx.device),
            C.byref(handle))
    if rc == -1:
        raise ValueError(ctx.lib.pdb_last_error(ctx.handle).decode())
    ctx._ok(rc, "pdb_matches_pack_colmap")
    return _native.Matches(ctx, handle, frames, int(counts.sum()))
