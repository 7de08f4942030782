"""Generate tests/golden/metrics.npz with the REFERENCE's own post-loop geometry (util/camera_transform.py, util/metric.py).

TEST INFRASTRUCTURE.  Run in the build container only:  python -m oracle.make_golden_metrics
pytorch3d is absent: `so3_relative_angle` and `get_world_to_view_transform` come from oracle/shims (restated from the published
pytorch3d semantics, version unpinned -- SURVEY 8c); everything else (`pose_encoding_to_camera`, `camera_to_rel_deg`,
`calculate_auc`, `calculate_auc_np`, `compute_ARE`) is the reference's unmodified code.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.ref_loader import REFERENCE_ROOT, load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "metrics.npz")
CASES = {"b2n8": (2, 8, 5), "b1n20": (1, 20, 6), "b3n3": (3, 3, 7)}  # name: (batch, frames, seed)


def random_poses(batch, frames, seed, near=None, noise=0.0):
    g = torch.Generator().manual_seed(seed)
    if near is not None:
        return near + noise * torch.randn(near.shape, generator=g)
    pose = torch.randn(batch, frames, 9, generator=g)
    pose[..., 7:] = 0.3 * pose[..., 7:]
    return pose


def main():
    torch.set_num_threads(1)
    ref = load_reference()
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "pose_diffusion"))
    from util import metric as ref_metric

    out = {}
    for name, (b, n, seed) in CASES.items():
        gt_pose = random_poses(b, n, seed)
        pred_pose = random_poses(b, n, seed + 100, near=gt_pose, noise=0.25)
        pred_pose[0, :2] = gt_pose[0, :2]  # pair (0,1) matches exactly: its rotation angle lands in the linear-extrapolation branch
        gt = ref.pose_encoding_to_camera(gt_pose)
        pred = ref.pose_encoding_to_camera(pred_pose)
        r_deg, t_deg = ref_metric.camera_to_rel_deg(pred, gt, torch.device("cpu"), b)
        out[f"{name}_gt_pose"], out[f"{name}_pred_pose"] = gt_pose.numpy(), pred_pose.numpy()
        out[f"{name}_R"], out[f"{name}_T"], out[f"{name}_fl"] = pred.R.numpy(), pred.T.numpy(), pred.focal_length.numpy()
        out[f"{name}_gt_R"], out[f"{name}_gt_T"] = gt.R.numpy(), gt.T.numpy()
        out[f"{name}_r_deg"], out[f"{name}_t_deg"] = r_deg.numpy(), t_deg.numpy()
        out[f"{name}_auc"] = np.float64(ref_metric.calculate_auc(r_deg, t_deg, max_threshold=30).item())
        out[f"{name}_auc_np"] = np.float64(ref_metric.calculate_auc_np(r_deg.numpy(), t_deg.numpy(), max_threshold=30))
        out[f"{name}_are"] = ref_metric.compute_ARE(pred.R, gt.R)
        print(name, r_deg.shape, float(r_deg.mean()), float(t_deg.mean()), out[f"{name}_auc"], out[f"{name}_auc_np"], float(out[f"{name}_are"].mean()))
    np.savez(OUT, **out)
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
