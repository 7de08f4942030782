"""Property-based check of the stream layouts on random match-set shapes: arbitrary pair order (repeated pairs, diagonal pairs,
both directions), segment sizes around the 32 / 64-row boundaries, any number of CTAs, the three streaming modes.  For every
generated case the kernel source -- run on the CPU emulation, tests/host/cuda_emu.h -- must give the same valid count and the
same gradient (up to summation order) for the paired layout as for the plain one, and the traversal replay must consume every
match exactly once."""
import ctypes

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import test_ggs_emulated_cpu as T
import test_layout_cpu as L
from posediffusion_b200 import synthetic as syn

SIZES = st.sampled_from([1, 2, 31, 32, 33, 63, 64, 65, 96, 127, 128, 129, 200])


@st.composite
def match_sets(draw):
    frames = draw(st.integers(2, 6))
    nseg = draw(st.integers(1, 12))
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    _, gt, start = syn.scene_matches(frames, 1, seed=int(rng.integers(1 << 20)))
    base, _, _ = syn.scene_matches(frames, 256, seed=int(rng.integers(1 << 20)))
    by_pair = {}
    for row, (a, b) in enumerate(base["i12"]):
        by_pair.setdefault((int(a), int(b)), []).append(row)
    kp1, kp2, i12 = [], [], []
    allow_diag = draw(st.booleans())
    for _ in range(nseg):
        a, b = (int(v) for v in rng.integers(0, frames, size=2))
        if a == b and not allow_diag:
            b = (a + 1) % frames
        count = draw(SIZES)
        if a != b:  # geometry-consistent rows of that ordered pair (so that many matches are valid)
            rows = rng.choice(by_pair[(a, b)], size=count, replace=True)
            kp1.append(base["kp1"][rows]); kp2.append(base["kp2"][rows])
        else:       # a diagonal pair: F' = 0 -> NaN errors, the reference's quirk
            kp1.append(rng.uniform(0, 224, size=(count, 2))); kp2.append(rng.uniform(0, 224, size=(count, 2)))
        i12.append(np.tile(np.array([[a, b]], dtype=np.int64), (count, 1)))
    m = {"kp1": np.concatenate(kp1), "kp2": np.concatenate(kp2), "i12": np.concatenate(i12), "img_shape": (frames, 3, 224, 224)}
    return m, start, draw(st.sampled_from(sorted(T.MODES))), draw(st.integers(1, 5))


@pytest.fixture(scope="module")
def libs():
    import __graft_entry__ as entry
    import os

    entry.build()
    emu = ctypes.CDLL(entry.build_emulator())
    emu.ggs_emu_run.restype = ctypes.c_int
    walk = ctypes.CDLL(os.path.join(T.ROOT, "build", "libgeom_host.so"))
    walk.ggs_host_walk.restype = ctypes.c_int
    walk.ggs_host_walk.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return emu, walk


@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow], derandomize=True)
@given(case=match_sets())
def test_paired_equals_plain_on_random_match_sets(libs, case):
    emu, walk = libs
    m, pose, mode, cpp = case
    for layout in ("plain", "paired"):
        rc, _, visits, _ = L._walk(walk, m, layout, {"resident": 0, "ring": 1, "stream": 2}[mode], cpp)
        assert rc == 0 and (visits == 1).all()
    out = {lay: T.run_kernel(emu, m, pose, lay, mode, cpp=cpp, eval_flags=(1, 1, 1)) for lay in ("plain", "paired")}
    a, b = out["plain"], out["paired"]
    assert a["scalars"][1] == b["scalars"][1]                      # valid counts: exact
    assert np.array_equal(np.isnan(a["grad"]), np.isnan(b["grad"]))  # NaN poisoning by diagonal pairs: same entries
    ok = ~np.isnan(a["grad"])
    if ok.any() and a["scalars"][1] > 0:
        gmax = np.abs(a["grad"][ok]).max() + 1e-30
        np.testing.assert_allclose(b["grad"][ok], a["grad"][ok], rtol=0, atol=1e-4 * gmax)
    assert np.array_equal(a["F"], b["F"])
