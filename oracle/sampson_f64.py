"""fp64 arbiters for the Sampson / GGS maths -- TEST INFRASTRUCTURE (see oracle/pose_oracle.py header).

The reference cannot run in fp64 (hard `.float()` casts at geometry_guided_sampling.py:167 and
embedding.py:31), so fp32-vs-fp32 disagreements between the CUDA path and the fp32 oracle are
arbitrated by:

  * `sampson_autograd_f64`   -- the reference's formula chain
                                evaluated in float64 with torch autograd;
  * `sampson_closed_form_f64`-- the two-stage closed form the CUDA kernels implement
                                (stage 1: per-pair 3x3 G = dL/dF'; stage 2: analytic adjoint of
                                pose -> F'), in numpy float64, independent of autograd.

Both follow util/geometry_guided_sampling.py:129-172, util/get_fundamental_matrix.py:14-51 and
util/camera_transform.py:85-97.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

LOG_FL_BIAS = 1.8
FL_MIN, FL_MAX = 0.1, 20.0


def sampson_autograd_f64(pose, matches: Dict, update_R=True, update_T=True, update_FL=True, sampson_max=10.0):
    """pose [N,9] (any float) -> dict(loss, n_valid, logged, grad [N,9], err [M]) in float64.
    Point coordinates are rounded to float32 first, as both fp32 implementations see them."""
    frames, _, height, width = matches["img_shape"]
    p = torch.as_tensor(np.asarray(pose), dtype=torch.float64).reshape(frames, 9).clone().requires_grad_(True)
    x1 = torch.from_numpy(np.asarray(matches["kp1"], dtype=np.float32).astype(np.float64))
    x2 = torch.from_numpy(np.asarray(matches["kp2"], dtype=np.float32).astype(np.float64))
    ia = torch.from_numpy(np.asarray(matches["i12"][:, 0], dtype=np.int64))
    ib = torch.from_numpy(np.asarray(matches["i12"][:, 1], dtype=np.int64))
    one = torch.ones(len(x1), 1, dtype=torch.float64)
    x1 = torch.cat([x1, one], 1)
    x2 = torch.cat([x2, one], 1)

    T, q, lam = p[:, :3], p[:, 3:7], p[:, 7:9]
    w, x, y, z = q.unbind(-1)
    s2 = 2.0 / (q * q).sum(-1)
    R = torch.stack(
        [
            1 - s2 * (y * y + z * z), s2 * (x * y - z * w), s2 * (x * z + y * w),
            s2 * (x * y + z * w), 1 - s2 * (x * x + z * z), s2 * (y * z - x * w),
            s2 * (x * z - y * w), s2 * (y * z + x * w), 1 - s2 * (x * x + y * y),
        ],
        -1,
    ).reshape(frames, 3, 3)
    focal = torch.clamp((lam + LOG_FL_BIAS).exp(), FL_MIN, FL_MAX).mean(0)
    if not update_R:
        R = R.detach()
    if not update_T:
        T = T.detach()
    if not update_FL:
        focal = focal.detach()
    D = torch.tensor([-1.0, -1.0, 1.0], dtype=torch.float64)
    Rcv = (R * D[None, None, :]).transpose(1, 2)
    tcv = T * D[None, :]
    scale = min(height, width) / 2.0
    Kinv = torch.zeros(3, 3, dtype=torch.float64)
    Kinv = Kinv + torch.diag(torch.stack([1 / (focal[0] * scale), 1 / (focal[1] * scale), torch.tensor(1.0, dtype=torch.float64)]))
    shift = torch.zeros(3, 3, dtype=torch.float64)
    shift[0, 2] = 1.0
    Kinv = Kinv - shift * (width / 2.0) / (focal[0] * scale)
    shift2 = torch.zeros(3, 3, dtype=torch.float64)
    shift2[1, 2] = 1.0
    Kinv = Kinv - shift2 * (height / 2.0) / (focal[1] * scale)

    def hat(v):
        o = torch.zeros_like(v[..., 0])
        return torch.stack([o, -v[..., 2], v[..., 1], v[..., 2], o, -v[..., 0], -v[..., 1], v[..., 0], o], -1).reshape(
            v.shape[:-1] + (3, 3)
        )

    R1, t1, R2, t2 = Rcv[ia], tcv[ia], Rcv[ib], tcv[ib]
    R12 = R2 @ R1.transpose(1, 2)
    t12 = t2 - (R12 @ t1[..., None])[..., 0]
    E = R12 @ hat(-(R12.transpose(1, 2) @ t12[..., None])[..., 0])
    Fm = (Kinv.T @ E @ Kinv).transpose(1, 2)  # F' = F^T : x1^T F' x2 = 0
    left = (x1[:, None, :] @ Fm)[:, 0, :]
    right = (Fm @ x2[:, :, None])[:, :, 0]
    top = (left * x2).sum(-1)
    bottom = left[:, 0] ** 2 + left[:, 1] ** 2 + right[:, 0] ** 2 + right[:, 1] ** 2
    err = top**2 / bottom
    keep = err < sampson_max
    n_valid = int(keep.sum())
    logged = float(torch.clamp(err.detach(), max=sampson_max).mean()) if len(err) else float("nan")
    if n_valid > 0:
        loss = err[keep].mean()
        loss.backward()
        grad = p.grad.numpy().copy()
        loss_v = float(loss)
    else:
        grad = np.full((frames, 9), np.nan)
        loss_v = float("nan")
    return {"loss": loss_v, "n_valid": n_valid, "logged": logged, "grad": grad, "err": err.detach().numpy()}


# ---------------------------------------------------------------------------------------------
# Closed form (what the CUDA kernels compute)
# ---------------------------------------------------------------------------------------------
def _hat_np(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def frame_terms(pose: np.ndarray, height: int, width: int):
    """Stage 0a: per frame R_cv = D R^T, t_cv = D T, A = hat(t_cv) R_cv; shared K^-1."""
    pose = np.asarray(pose, dtype=np.float64)
    n = pose.shape[0]
    D = np.diag([-1.0, -1.0, 1.0])
    Rp = np.zeros((n, 3, 3))
    for i in range(n):
        w, x, y, z = pose[i, 3:7]
        s2 = 2.0 / (w * w + x * x + y * y + z * z)
        Rp[i] = np.array(
            [
                [1 - s2 * (y * y + z * z), s2 * (x * y - z * w), s2 * (x * z + y * w)],
                [s2 * (x * y + z * w), 1 - s2 * (x * x + z * z), s2 * (y * z - x * w)],
                [s2 * (x * z - y * w), s2 * (y * z + x * w), 1 - s2 * (x * x + y * y)],
            ]
        )
    Rcv = np.einsum("ij,nkj->nik", D, Rp)
    tcv = pose[:, :3] @ D
    A = np.stack([_hat_np(tcv[i]) @ Rcv[i] for i in range(n)])
    raw = np.exp(pose[:, 7:9] + LOG_FL_BIAS)
    fl = np.clip(raw, FL_MIN, FL_MAX)
    in_range = ((raw >= FL_MIN) & (raw <= FL_MAX)).astype(np.float64)
    scale = min(height, width) / 2.0
    fpx = fl.mean(0) * scale
    Kinv = np.array([[1 / fpx[0], 0, -(width / 2.0) / fpx[0]], [0, 1 / fpx[1], -(height / 2.0) / fpx[1]], [0, 0, 1.0]])
    return Rp, Rcv, tcv, A, fl, in_range, fpx, Kinv, scale


def pair_F(Rcv, A, Kinv, a: int, b: int) -> np.ndarray:
    """Stage 0b: F'_{ab} = K^-T M K^-1 with M = E^T = -(A_a R_b^T + R_a A_b^T).
    A diagonal pair (a == b) has E = 0 analytically; the reference's fp32 chain yields exactly 0 there
    (t12 = t - (R R^T) t rounds to 0), hence 0/0 = NaN errors.  We make that explicit."""
    return Kinv.T @ pair_M(Rcv, A, a, b) @ Kinv


def pair_M(Rcv, A, a: int, b: int) -> np.ndarray:
    if a == b:
        return np.zeros((3, 3))
    return -(A[a] @ Rcv[b].T + Rcv[a] @ A[b].T)


def sampson_closed_form_f64(pose, matches: Dict, update_R=True, update_T=True, update_FL=True, sampson_max=10.0):
    """Returns dict(loss, n_valid, logged, grad [N,9], G {pair: 3x3}, F {pair: 3x3})."""
    frames, _, height, width = matches["img_shape"]
    pose = np.asarray(pose, dtype=np.float64).reshape(frames, 9)
    Rp, Rcv, tcv, A, fl, in_range, fpx, Kinv, scale = frame_terms(pose, height, width)
    kp1 = np.asarray(matches["kp1"], dtype=np.float32).astype(np.float64)
    kp2 = np.asarray(matches["kp2"], dtype=np.float32).astype(np.float64)
    ia = np.asarray(matches["i12"][:, 0], dtype=np.int64)
    ib = np.asarray(matches["i12"][:, 1], dtype=np.int64)
    pair = ia * frames + ib

    G: Dict[int, np.ndarray] = {}
    Fs: Dict[int, np.ndarray] = {}
    n_valid = 0
    clamp_sum = 0.0
    loss_sum = 0.0
    # ---- stage 1: per match error + per-pair gradient wrt F' ----
    for pid in np.unique(pair):
        a, b = divmod(int(pid), frames)
        Fp = pair_F(Rcv, A, Kinv, a, b)
        Fs[int(pid)] = Fp
        sel = pair == pid
        x1 = np.concatenate([kp1[sel], np.ones((sel.sum(), 1))], 1)
        x2 = np.concatenate([kp2[sel], np.ones((sel.sum(), 1))], 1)
        left = x1 @ Fp
        right = x2 @ Fp.T
        top = (left * x2).sum(1)
        bottom = left[:, 0] ** 2 + left[:, 1] ** 2 + right[:, 0] ** 2 + right[:, 1] ** 2
        with np.errstate(invalid="ignore", divide="ignore"):
            err = top**2 / bottom
        keep = err < sampson_max  # NaN (diagonal pairs, F'=0) fails the test
        clamp_sum += np.where(err > sampson_max, sampson_max, err).sum()  # NaN propagates like torch.clamp
        n_valid += int(keep.sum())
        loss_sum += err[keep].sum()
        # validity enters as a 0/1 weight (not a select): this mirrors the reference's autograd, where a
        # filtered-out match back-propagates 0 * d(top^2/bottom) -- NaN when bottom == 0 (diagonal pair).
        wgt = keep.astype(np.float64)
        with np.errstate(invalid="ignore", divide="ignore"):
            ca = wgt * (2 * top / bottom)
            cb = wgt * (2 * err / bottom)
        lz = left.copy()
        lz[:, 2] = 0
        rz = right.copy()
        rz[:, 2] = 0
        with np.errstate(invalid="ignore"):
            G[int(pid)] = (
                np.einsum("m,mi,mj->ij", ca, x1, x2)
                - np.einsum("m,mi,mj->ij", cb, x1, lz)
                - np.einsum("m,mi,mj->ij", cb, rz, x2)
            )
    # ---- stage 2: adjoint of pose -> F' ----
    gR = np.zeros((frames, 3, 3))
    gA = np.zeros((frames, 3, 3))
    gK = np.zeros((3, 3))
    for pid, Gp in G.items():
        a, b = divmod(pid, frames)
        M = pair_M(Rcv, A, a, b)
        H = Kinv @ Gp @ Kinv.T
        gK += M @ Kinv @ Gp.T + M.T @ Kinv @ Gp
        gA[a] += -H @ Rcv[b]
        gR[b] += -H.T @ A[a]
        gR[a] += -H @ A[b]
        gA[b] += -H.T @ Rcv[a]
    grad = np.zeros((frames, 9))
    D = np.diag([-1.0, -1.0, 1.0])
    g_fpx = np.array(
        [
            (-gK[0, 0] + (width / 2.0) * gK[0, 2]) / fpx[0] ** 2,
            (-gK[1, 1] + (height / 2.0) * gK[1, 2]) / fpx[1] ** 2,
        ]
    )
    for i in range(frames):
        gRcv = gR[i] - _hat_np(tcv[i]) @ gA[i]
        Wm = gA[i] @ Rcv[i].T
        gt = np.array([Wm[2, 1] - Wm[1, 2], Wm[0, 2] - Wm[2, 0], Wm[1, 0] - Wm[0, 1]])
        if update_T:
            grad[i, :3] = D @ gt
        if update_R:
            gRp = (D @ gRcv).T
            w, x, y, z = pose[i, 3:7]
            q = np.array([w, x, y, z])
            s2 = 2.0 / (q @ q)
            B = np.array(
                [
                    [-(y * y + z * z), x * y - z * w, x * z + y * w],
                    [x * y + z * w, -(x * x + z * z), y * z - x * w],
                    [x * z - y * w, y * z + x * w, -(x * x + y * y)],
                ]
            )
            dB = [
                np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]]),
                np.array([[0, y, z], [y, -2 * x, -w], [z, w, -2 * x]]),
                np.array([[-2 * y, x, w], [x, 0, z], [-w, z, -2 * y]]),
                np.array([[-2 * z, -w, x], [w, -2 * z, y], [x, y, 0]]),
            ]
            gB = (gRp * B).sum()
            for k in range(4):
                grad[i, 3 + k] = -s2 * s2 * q[k] * gB + s2 * (gRp * dB[k]).sum()
        if update_FL:
            grad[i, 7:9] = g_fpx * scale / frames * fl[i] * in_range[i]
    total = len(pair)
    out = {
        "n_valid": n_valid,
        "logged": clamp_sum / total if total else float("nan"),
        "G": G,
        "F": Fs,
    }
    if n_valid > 0:
        out["loss"] = loss_sum / n_valid
        out["grad"] = grad / n_valid
    else:
        out["loss"] = float("nan")
        out["grad"] = np.full((frames, 9), np.nan)
    return out


# ---------------------------------------------------------------------------------------------
# Same closed form, vectorised for BASELINE-size match sets (778 240 / 25 886 720 rows)
# ---------------------------------------------------------------------------------------------
def sampson_closed_form_f64_large(pose, matches: Dict, update_R=True, update_T=True, update_FL=True, sampson_max=10.0,
                                  chunk_rows: int = 1 << 20):
    """`sampson_closed_form_f64` for pair-contiguous match sets of any size: identical maths (stage 1 per match in float64
    on the float32-rounded coordinates, stage 2 per pair segment), evaluated in row chunks with `np.add.reduceat` over the
    maximal runs of equal (i12[:,0], i12[:,1]) instead of one boolean mask per pair.  Returns dict(loss, n_valid, logged,
    grad [N,9]); additionally `band`: the number of matches whose error lies within 1e-5 (relative) of the threshold,
    i.e. the matches whose validity an fp32 evaluation may legitimately flip."""
    frames, _, height, width = matches["img_shape"]
    pose = np.asarray(pose, dtype=np.float64).reshape(frames, 9)
    Rp, Rcv, tcv, A, fl, in_range, fpx, Kinv, scale = frame_terms(pose, height, width)
    ia = np.asarray(matches["i12"][:, 0], dtype=np.int64)
    ib = np.asarray(matches["i12"][:, 1], dtype=np.int64)
    total = len(ia)
    pair = ia * frames + ib
    # maximal runs of equal pair index (the segments of the device layout)
    starts = np.flatnonzero(np.concatenate([[True], pair[1:] != pair[:-1]])) if total else np.zeros(0, np.int64)
    seg_pair = pair[starts] if total else np.zeros(0, np.int64)
    # F' of every pair that occurs
    uniq = np.unique(seg_pair)
    F_of = {}
    for pid in uniq:
        a, b = divmod(int(pid), frames)
        F_of[int(pid)] = pair_F(Rcv, A, Kinv, a, b)
    F_seg = np.stack([F_of[int(p)] for p in seg_pair]) if len(seg_pair) else np.zeros((0, 3, 3))
    G_seg = np.zeros((len(starts), 3, 3))
    n_valid = 0
    band = 0
    clamp_sum = 0.0
    loss_sum = 0.0
    seg_of_row_start = np.searchsorted(starts, np.arange(0, total, chunk_rows), side="right") - 1 if total else []
    for ci, r0 in enumerate(range(0, total, chunk_rows)):
        r1 = min(total, r0 + chunk_rows)
        s0 = int(seg_of_row_start[ci])
        s1 = int(np.searchsorted(starts, r1 - 1, side="right"))  # segments [s0, s1) intersect the chunk
        local_starts = np.maximum(starts[s0:s1], r0) - r0
        seg_id = np.repeat(np.arange(s0, s1), np.diff(np.concatenate([local_starts, [r1 - r0]])))
        kp1 = np.asarray(matches["kp1"][r0:r1], dtype=np.float32).astype(np.float64)
        kp2 = np.asarray(matches["kp2"][r0:r1], dtype=np.float32).astype(np.float64)
        n = r1 - r0
        x1 = np.concatenate([kp1, np.ones((n, 1))], 1)
        x2 = np.concatenate([kp2, np.ones((n, 1))], 1)
        Fm = F_seg[seg_id]                                   # [n,3,3]
        left = np.einsum("mi,mij->mj", x1, Fm)
        right = np.einsum("mij,mj->mi", Fm, x2)
        top = (left * x2).sum(1)
        bottom = left[:, 0] ** 2 + left[:, 1] ** 2 + right[:, 0] ** 2 + right[:, 1] ** 2
        with np.errstate(invalid="ignore", divide="ignore"):
            err = top**2 / bottom
            keep = err < sampson_max
            band += int((np.abs(err - sampson_max) <= 1e-5 * sampson_max).sum())
            clamp_sum += np.where(err > sampson_max, sampson_max, err).sum()
            n_valid += int(keep.sum())
            loss_sum += err[keep].sum()
            wgt = keep.astype(np.float64)
            ca = wgt * (2 * top / bottom)
            cb = wgt * (2 * err / bottom)
            lz = left.copy()
            lz[:, 2] = 0
            rz = right.copy()
            rz[:, 2] = 0
            # G_ij = sum ca x1_i x2_j - cb x1_i lz_j - cb rz_i x2_j   per match, then summed per segment
            w = ca[:, None] * x2 - cb[:, None] * lz          # [n,3]
            per = x1[:, :, None] * w[:, None, :] - (cb[:, None] * rz)[:, :, None] * x2[:, None, :]
            G_seg[s0:s1] += np.add.reduceat(per.reshape(n, 9), local_starts, axis=0).reshape(-1, 3, 3)
    # ---- stage 2: adjoint of pose -> F' (as above) ----
    gR = np.zeros((frames, 3, 3))
    gA = np.zeros((frames, 3, 3))
    gK = np.zeros((3, 3))
    for s, pid in enumerate(seg_pair):
        a, b = divmod(int(pid), frames)
        Gp = G_seg[s]
        M = pair_M(Rcv, A, a, b)
        H = Kinv @ Gp @ Kinv.T
        gK += M @ Kinv @ Gp.T + M.T @ Kinv @ Gp
        gA[a] += -H @ Rcv[b]
        gR[b] += -H.T @ A[a]
        gR[a] += -H @ A[b]
        gA[b] += -H.T @ Rcv[a]
    grad = np.zeros((frames, 9))
    D = np.diag([-1.0, -1.0, 1.0])
    g_fpx = np.array([(-gK[0, 0] + (width / 2.0) * gK[0, 2]) / fpx[0] ** 2, (-gK[1, 1] + (height / 2.0) * gK[1, 2]) / fpx[1] ** 2])
    for i in range(frames):
        gRcv = gR[i] - _hat_np(tcv[i]) @ gA[i]
        Wm = gA[i] @ Rcv[i].T
        gt = np.array([Wm[2, 1] - Wm[1, 2], Wm[0, 2] - Wm[2, 0], Wm[1, 0] - Wm[0, 1]])
        if update_T:
            grad[i, :3] = D @ gt
        if update_R:
            gRp = (D @ gRcv).T
            w, x, y, z = pose[i, 3:7]
            q = np.array([w, x, y, z])
            s2 = 2.0 / (q @ q)
            B = np.array([[-(y * y + z * z), x * y - z * w, x * z + y * w], [x * y + z * w, -(x * x + z * z), y * z - x * w],
                          [x * z - y * w, y * z + x * w, -(x * x + y * y)]])
            dB = [np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]]), np.array([[0, y, z], [y, -2 * x, -w], [z, w, -2 * x]]),
                  np.array([[-2 * y, x, w], [x, 0, z], [-w, z, -2 * y]]), np.array([[-2 * z, -w, x], [w, -2 * z, y], [x, y, 0]])]
            gB = (gRp * B).sum()
            for k in range(4):
                grad[i, 3 + k] = -s2 * s2 * q[k] * gB + s2 * (gRp * dB[k]).sum()
        if update_FL:
            grad[i, 7:9] = g_fpx * scale / frames * fl[i] * in_range[i]
    out = {"n_valid": n_valid, "band": band, "logged": clamp_sum / total if total else float("nan")}
    if n_valid > 0:
        out["loss"] = loss_sum / n_valid
        out["grad"] = grad / n_valid
    else:
        out["loss"] = float("nan")
        out["grad"] = np.full((frames, 9), np.nan)
    return out
