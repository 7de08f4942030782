"""CPU-only checks of the camera alignment (demo.py:126-128 -> pytorch3d corresponding_cameras_alignment):

* the oracle restatement (oracle/cameras_alignment.py; pytorch3d is absent and unpinned -> parity UNPINNED, stated there)
  satisfies the property that defines the operation: source cameras that differ from the targets by a world similarity
  transform come back onto the targets;
* the device maths of csrc/align.cuh, compiled for the host by tests/host/geom_host.cu and run exactly as the two kernels run
  it, agrees with the oracle; its 3x3 `V U^T` agrees with numpy's SVD, including sign / ordering invariance.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import cameras_alignment as oca


@pytest.fixture(scope="module")
def harness():
    import __graft_entry__ as entry

    entry.build_host_harness()
    return ctypes.CDLL(os.path.join(ROOT, "build", "libgeom_host.so"))


def random_rotations(n, rng):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
                     2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).reshape(n, 3, 3)


def similarity_scene(n, seed, noise=0.0):
    """Target cameras and source cameras = the targets seen from a world moved by X' = (X R_w + t_w) / s (row vectors)."""
    rng = np.random.default_rng(seed)
    R_tgt, T_tgt = random_rotations(n, rng), rng.normal(size=(n, 3))
    R_w, t_w, s = random_rotations(1, rng)[0], rng.normal(size=3), float(rng.uniform(0.5, 2.0))
    # choose the source so that align_R R_src = R_tgt and align_T R_src + s T_src = T_tgt with align_R = R_w, align_T = t_w
    R_src = R_w.T @ R_tgt
    T_src = (T_tgt - t_w @ R_src) / s
    R_src = R_src + noise * rng.normal(size=R_src.shape)
    T_src = T_src + noise * rng.normal(size=T_src.shape)
    return R_src, T_src, R_tgt, T_tgt, (R_w, t_w, s)


def as_t(*arrays):
    return [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)) for a in arrays]


@pytest.mark.parametrize("n", [2, 5, 20, 80])
def test_oracle_recovers_a_world_similarity(n):
    R_src, T_src, R_tgt, T_tgt, (R_w, t_w, s) = similarity_scene(n, seed=n)
    a_R, a_T, a_s = oca.align_camera_extrinsics(*[t.double() for t in as_t(R_src, T_src, R_tgt, T_tgt)])
    np.testing.assert_allclose(a_R.numpy(), R_w, atol=1e-5)
    np.testing.assert_allclose(a_T.numpy(), t_w, atol=1e-5)
    np.testing.assert_allclose(float(a_s), s, rtol=1e-5)
    R_new, T_new = oca.corresponding_cameras_alignment(*[t.double() for t in as_t(R_src, T_src, R_tgt, T_tgt)])
    np.testing.assert_allclose(R_new.numpy(), R_tgt, atol=1e-5)
    np.testing.assert_allclose(T_new.numpy(), T_tgt, atol=1e-5)


def _host_align(harness, R_src, T_src, R_tgt, T_tgt, estimate_scale=True, eps=1e-9):
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (R_src, T_src, R_tgt, T_tgt)]
    n = len(arrs[0])
    Ro, To, al = np.zeros((n, 3, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros(13, np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    harness.cameras_align_host(*[P(a) for a in arrs], n, int(estimate_scale), ctypes.c_float(eps), P(Ro), P(To), P(al))
    return Ro, To, al


@pytest.mark.parametrize("n,noise,estimate_scale", [(1, 0.0, True), (2, 0.0, True), (5, 0.05, True), (20, 0.2, True), (20, 0.2, False), (80, 0.02, True)])
def test_device_maths_on_host_matches_oracle(harness, n, noise, estimate_scale):
    R_src, T_src, R_tgt, T_tgt, _ = similarity_scene(n, seed=100 + n, noise=noise)
    Ro, To, al = _host_align(harness, R_src, T_src, R_tgt, T_tgt, estimate_scale)
    want_R, want_T = oca.corresponding_cameras_alignment(*as_t(R_src, T_src, R_tgt, T_tgt), estimate_scale=estimate_scale)
    a_R, a_T, a_s = oca.align_camera_extrinsics(*as_t(R_src, T_src, R_tgt, T_tgt), estimate_scale=estimate_scale)
    np.testing.assert_allclose(al[:9].reshape(3, 3), a_R.numpy(), atol=5e-6)
    np.testing.assert_allclose(al[12], float(a_s), rtol=2e-5)
    np.testing.assert_allclose(al[9:12], a_T.numpy(), atol=2e-5 * max(1.0, np.abs(a_T.numpy()).max()))
    np.testing.assert_allclose(Ro, want_R.numpy(), atol=1e-5)
    np.testing.assert_allclose(To, want_T.numpy(), atol=3e-5 * max(1.0, np.abs(want_T.numpy()).max()))


def test_v_ut_matches_numpy_svd_and_is_a_rotation_for_mixed_inputs(harness):
    rng = np.random.default_rng(0)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    cases = [rng.normal(size=(3, 3)) for _ in range(50)]
    cases += [random_rotations(1, rng)[0] * rng.uniform(0.1, 1.0) for _ in range(10)]            # mean of near-identical rotations
    cases += [np.diag([1.0, 1.0, -1.0]) @ random_rotations(1, rng)[0] for _ in range(5)]        # reflections: det < 0
    cases += [np.mean(random_rotations(4, rng), 0) for _ in range(20)]                            # what RRcov looks like
    for M in cases:
        M32 = np.ascontiguousarray(M, dtype=np.float32)
        out = np.zeros((3, 3), np.float32)
        harness.svd3_v_ut_host(P(M32), P(out))
        U, S, Vh = np.linalg.svd(M32.astype(np.float64))
        want = Vh.T @ U.T
        cond = S[0] / max(S[-1], 1e-30)
        np.testing.assert_allclose(out, want, atol=2e-6 * max(1.0, cond))
        np.testing.assert_allclose(out @ out.T, np.eye(3), atol=1e-5)
    # rank-deficient input: the completed basis keeps det(V U^T) = +1 and reproduces the well-defined part
    M = np.outer([1.0, 2.0, 2.0], [0.0, 3.0, 4.0]) + np.outer([2.0, -1.0, 0.0], [1.0, 0.0, 0.0])
    out = np.zeros((3, 3), np.float32)
    harness.svd3_v_ut_host(P(np.ascontiguousarray(M, dtype=np.float32)), P(out))
    assert abs(np.linalg.det(out.astype(np.float64)) - 1.0) < 1e-5
    U, S, Vh = np.linalg.svd(M)
    np.testing.assert_allclose(out.astype(np.float64) @ U[:, :2], Vh[:2].T, atol=1e-5)  # V U^T maps u_i -> v_i on the range


def test_python_mirror_contract():
    import posediffusion_b200 as pdb
    from posediffusion_b200 import _native

    cams = pdb.PerspectiveCameras(focal_length=torch.ones(3, 2), R=torch.eye(3).repeat(3, 1, 1), T=torch.zeros(3, 3))
    other = pdb.PerspectiveCameras(focal_length=torch.ones(2, 2), R=torch.eye(3).repeat(2, 1, 1), T=torch.zeros(2, 3))
    with pytest.raises(ValueError):
        pdb.corresponding_cameras_alignment(cams, other)
    with pytest.raises(ValueError):
        pdb.corresponding_cameras_alignment(cams, cams, mode="other")
    with pytest.raises(_native.NativeError):
        pdb.corresponding_cameras_alignment(cams, cams)  # CPU tensors: no fallback


@pytest.mark.parametrize("n,noise,estimate_scale", [(1, 0.0, True), (2, 0.0, True), (20, 0.2, True), (20, 0.2, False), (80, 0.02, True), (200, 0.1, True)])
def test_kernel_bodies_on_the_emulated_grid_match_oracle(n, noise, estimate_scale):
    """The two kernels of pdb_cameras_align (one warp with shuffles for the estimate, 128-thread CTAs for the application) run on
    the CPU emulation of the execution model (tests/host/cuda_emu.h) with the launch geometry of the C entry point."""
    import __graft_entry__ as entry

    lib = ctypes.CDLL(entry.build_emulator())
    R_src, T_src, R_tgt, T_tgt, _ = similarity_scene(n, seed=300 + n, noise=noise)
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (R_src, T_src, R_tgt, T_tgt)]
    Ro, To, al = np.zeros((n, 3, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros(13, np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.cameras_align_emu(*[P(a) for a in arrs], n, int(estimate_scale), ctypes.c_float(1e-9), P(Ro), P(To), P(al))
    want_R, want_T = oca.corresponding_cameras_alignment(*as_t(R_src, T_src, R_tgt, T_tgt), estimate_scale=estimate_scale)
    np.testing.assert_allclose(Ro, want_R.numpy(), atol=1e-5)
    np.testing.assert_allclose(To, want_T.numpy(), atol=3e-5 * max(1.0, np.abs(want_T.numpy()).max()))
