"""CPU execution of the fp32 DENOISER KERNEL SOURCE (csrc/denoiser.cuh: the persistent cooperative kernel that runs the 43
stages of a diffusion step and the fused DDPM update; both stage hand-overs: flag-carrying activation words -- the buffers
start out full of stale-version NaN words, so accepting a wrong version fails the comparison -- and group barriers) through the execution-model emulation of
tests/host/cuda_emu.h, against the fixtures produced by the reference's own `Denoiser` / `GaussianDiffusion` modules.
Weights are re-laid out exactly as pdb_denoiser_load does it (restated as host loops in tests/host/kernels_emu.cpp).
Same tolerances as the GPU parity tests.  Test infrastructure: kernel logic only, see tests/test_ggs_emulated_cpu.py."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden
from oracle import pose_oracle as po
from posediffusion_b200 import _native
from posediffusion_b200 import synthetic as syn


@pytest.fixture(scope="module")
def emu():
    import __graft_entry__ as entry

    entry.build()
    lib = C.CDLL(entry.build_emulator())
    lib.denoiser_emu_run.restype = C.c_int
    return lib


@pytest.fixture(autouse=True, params=["flags", "barriers"])
def handover(request, monkeypatch):
    monkeypatch.setenv("PDB_DEN_FLAG", "1" if request.param == "flags" else "0")  # read by tests/host/kernels_emu.cpp per run
    return request.param


@pytest.fixture(scope="module")
def weights():
    g = load_golden("denoiser.npz")
    state = syn.random_denoiser_state(int(g["weight_seed"]), float(g["bias_std"]))
    tensors = [np.ascontiguousarray(state[name].numpy(), dtype=np.float32) for name in syn.denoiser_param_shapes()]
    assert len(tensors) == _native.PDB_NUM_WEIGHT_TENSORS
    return tensors


def run_steps(lib, tensors, x, z, t_hi, t_lo, draws=None, guide_below=0, grid=3, token_tile=None):
    B, N, _ = x.shape
    S = B * N
    if token_tile is None:  # pick_token_tile of csrc/api_sampler.cu
        token_tile = min((8, 16, 20, 24, 32), key=lambda c: ((S + c - 1) // c * c, -c))
    arr = (C.c_void_p * len(tensors))(*[t.ctypes.data for t in tensors])
    sched = _native.schedule_table()
    xs = np.ascontiguousarray(x, np.float32).reshape(S, 9).copy()
    zs = np.ascontiguousarray(z, np.float32).reshape(S, 384)
    eps, x0, mean = (np.zeros((S, 9), np.float32) for _ in range(3))
    trail = np.zeros((101, S, 9), np.float32)
    P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    d = None if draws is None else np.ascontiguousarray(draws, np.float32)
    rc = lib.denoiser_emu_run(arr, P(sched), B, N, t_hi, t_lo, guide_below, P(xs), P(zs), P(d), P(trail), P(eps), P(x0), P(mean), grid, token_tile)
    assert rc == 0
    return dict(x=xs.reshape(B, N, 9), eps=eps.reshape(B, N, 9), x0=x0.reshape(B, N, 9), mean=mean.reshape(B, N, 9), trail=trail.reshape(101, B, N, 9))


@pytest.mark.parametrize("tag,grid", [("b1n5", 3), ("b1n20", 4), ("b2n20", 6)])
def test_emulated_denoiser_forward_vs_reference(emu, weights, tag, grid):
    """Denoiser.forward (models/denoiser.py:53-76) = one step of the kernel; eps within 3e-5 of the reference's output."""
    g = load_golden("denoiser.npz")
    r = run_steps(emu, weights, g[f"{tag}_x"], g[f"{tag}_z"], int(g[f"{tag}_t"]), int(g[f"{tag}_t"]), grid=grid)
    np.testing.assert_allclose(r["eps"], g[f"{tag}_eps"], rtol=0, atol=3e-5)


@pytest.mark.parametrize("t", [99, 11, 0])
def test_emulated_p_sample_teacher_forced_vs_reference(emu, weights, t):
    """p_sample (gaussian_diffuser.py:249-282) = denoiser + x0 + posterior mean + noise, fused in the kernel's tail stage."""
    g = load_golden("p_sample.npz")
    draws = np.zeros((101, 20, 9), np.float32)
    draws[1 + (99 - t)] = g[f"t{t}_noise"].reshape(20, 9)
    r = run_steps(emu, weights, g[f"t{t}_x"], g["z"], t, t, draws=draws, grid=3)
    scale = np.abs(g[f"t{t}_x0"]).max()
    b_t = abs(float(po.diffusion_schedule()["sqrt_recipm1_alphas_cumprod"][t]))
    np.testing.assert_allclose(r["x0"], g[f"t{t}_x0"], rtol=0, atol=1e-5 * scale + 3e-5 * b_t)
    np.testing.assert_allclose(r["x"], g[f"t{t}_pred"], rtol=0, atol=2e-5 * np.abs(g[f"t{t}_pred"]).max() + 1e-5)


def test_emulated_multi_step_launch_equals_single_steps(emu, weights):
    """One launch over steps 99..98 (persistent loop, z-projection hoisted) == two single-step launches fed with its own
    trajectory; the grid size does not change the result beyond summation order (none here: every output has one owner)."""
    g = load_golden("p_sample.npz")
    rng = np.random.default_rng(0)
    draws = rng.normal(size=(101, 20, 9)).astype(np.float32)
    x = g["t99_x"]
    fused = run_steps(emu, weights, x, g["z"], 99, 98, draws=draws, grid=4)
    cur = x
    for k, t in enumerate((99, 98)):
        one = run_steps(emu, weights, cur, g["z"], t, t, draws=draws, grid=3)
        np.testing.assert_array_equal(one["x"], fused["trail"][1 + k])
        cur = one["x"]
    np.testing.assert_array_equal(fused["x"], cur)


def test_emulated_sequence_spanning_token_tiles(emu, weights):
    """A sequence cut into several token tiles (20 frames, tiles of 8): attention items read key / value rows that other tiles'
    items own (the case the flag-carrying qkv buffer is double buffered for); two steps in one launch, odd grid sizes."""
    g = load_golden("p_sample.npz")
    rng = np.random.default_rng(1)
    draws = rng.normal(size=(101, 20, 9)).astype(np.float32)
    want = run_steps(emu, weights, g["t99_x"], g["z"], 99, 98, draws=draws, grid=4, token_tile=20)
    for grid in (5, 7):
        got = run_steps(emu, weights, g["t99_x"], g["z"], 99, 98, draws=draws, grid=grid, token_tile=8)
        np.testing.assert_allclose(got["x"], want["x"], rtol=0, atol=2e-5 * np.abs(want["x"]).max())
        np.testing.assert_allclose(got["eps"], want["eps"], rtol=0, atol=3e-5)
