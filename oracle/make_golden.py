"""Generate tests/golden/*.npz by running the REFERENCE's own modules (imported from /root/reference).

TEST INFRASTRUCTURE.  Run in the build container only:  python -m oracle.make_golden
The reference has no tests or golden vectors of its own for this path, so these
fixtures ARE the pin: every array below is produced by unmodified reference code
(models/denoiser.py, models/gaussian_diffuser.py, util/geometry_guided_sampling.py, ...) behind
the pytorch3d/hydra shims in oracle/shims.  Inputs are regenerated from seeds by
`posediffusion_b200.synthetic`; a weight checksum is stored to detect generator drift.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.ref_loader import build_reference_sampler, load_reference  # noqa: E402
from posediffusion_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 7
BIAS_STD = 0.05


def weight_checksum(state) -> float:
    return float(sum(v.double().abs().sum() for v in state.values()))


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = fn(*a, **k)
    return out, buf.getvalue()


def golden_schedule(ref):
    dif = ref.GaussianDiffusion(beta_schedule="custom")
    np.savez(os.path.join(OUT, "schedule.npz"), **{k: v.numpy() for k, v in dif.state_dict().items()})


def golden_denoiser(ref, sampler, state):
    out = {"weight_checksum": weight_checksum(state), "weight_seed": WEIGHT_SEED, "bias_std": BIAS_STD}
    gen = torch.Generator().manual_seed(11)
    for tag, (b, n, t) in {"b1n5": (1, 5, 99), "b1n20": (1, 20, 37), "b2n20": (2, 20, 3), "b1n80": (1, 80, 0)}.items():
        x = torch.randn(b, n, 9, generator=gen) * 2.0
        z = torch.randn(b, n, 384, generator=gen)
        tt = torch.full((b,), t, dtype=torch.long)
        with torch.no_grad():
            eps = sampler.model(x, tt, z)
        out[f"{tag}_x"], out[f"{tag}_z"], out[f"{tag}_t"], out[f"{tag}_eps"] = x.numpy(), z.numpy(), np.int64(t), eps.numpy()
    np.savez(os.path.join(OUT, "denoiser.npz"), **out)


def golden_p_sample(ref, sampler):
    """Teacher-forced single steps.  torch.randn_like is patched to return our injected draw
    (the reference draws on its device generator, gaussian_diffuser.py:278)."""
    out = {}
    gen = torch.Generator().manual_seed(12)
    n = 20
    z = torch.randn(1, n, 384, generator=gen)
    out["z"] = z.numpy()
    for t in (99, 50, 11, 10, 9, 0):
        x = torch.randn(1, n, 9, generator=gen)
        noise = torch.randn(1, n, 9, generator=gen)
        orig = torch.randn_like
        torch.randn_like = lambda *_a, **_k: noise.clone()
        try:
            pred, x0 = sampler.p_sample(x=x, t=t, z=z)
        finally:
            torch.randn_like = orig
        out[f"t{t}_x"], out[f"t{t}_noise"], out[f"t{t}_pred"], out[f"t{t}_x0"] = x.numpy(), noise.numpy(), pred.numpy(), x0.numpy()
    np.savez(os.path.join(OUT, "p_sample.npz"), **out)


def _processed(matches):
    """Replicates the dict the reference builds at geometry_guided_sampling.py:16-45 by calling it
    with a patched GGS_optimize that captures its `processed_matches` argument."""
    import util.geometry_guided_sampling as mod

    grabbed = {}

    def capture(model_mean, t, processed, **kw):
        grabbed["p"] = processed
        return model_mean

    orig = mod.GGS_optimize
    mod.GGS_optimize = capture
    try:
        mod.geometry_guided_sampling(torch.zeros(1, matches["img_shape"][0], 9), 0, matches, syn.default_ggs_cfg())
    finally:
        mod.GGS_optimize = orig
    return grabbed["p"]


def sampson_case(ref, pose, matches, flags):
    proc = _processed(matches)
    p = torch.from_numpy(pose)[None].clone().requires_grad_(True)
    with torch.enable_grad():
        valid, logged = ref.compute_sampson_distance(
            p, 0, proc, update_R=flags[0], update_T=flags[1], update_FL=flags[2], sampson_max=10
        )
        n_valid = len(valid)
        if n_valid:
            valid.mean().backward()
            grad = p.grad[0].numpy().copy()
            loss = float(valid.mean())
        else:
            grad, loss = np.full_like(pose, np.nan), float("nan")
    return dict(n_valid=np.int64(n_valid), logged=np.float32(logged), loss=np.float32(loss), grad=grad)


def golden_sampson(ref):
    out = {}
    cases = {}
    m, gt, start = syn.scene_matches(6, 64, seed=21)
    cases["scene6"] = (start, m)
    m, gt, start = syn.scene_matches(5, 24, seed=22, ordered=False, ragged=True)
    cases["ragged5"] = (start, m)
    # uniform-random correspondences (the bench workload): only a small fraction passes s < 10
    _, _, start = syn.scene_matches(5, 4, seed=23)
    cases["uniform5"] = (start, syn.uniform_matches(5, 400, seed=23))
    # no valid match at all (mean over an empty set)
    cases["empty5"] = (np.random.default_rng(23).normal(size=(5, 9)).astype(np.float32), syn.uniform_matches(5, 400, seed=23))
    # diagonal pair present (F = 0 -> NaN error -> dropped, logged value NaN-poisoned)
    m, gt, start = syn.scene_matches(4, 16, seed=24)
    m["i12"][:8] = np.array([[1, 1]])
    cases["diag4"] = (start, m)
    # saturated focal clamp on one frame (gradient mask / clamp pass-through)
    m, gt, start = syn.scene_matches(4, 32, seed=25)
    start = start.copy()
    start[2, 7:9] = 5.0
    cases["clamp4"] = (start, m)
    for tag, (pose, matches) in cases.items():
        out[f"{tag}_pose"] = pose
        out[f"{tag}_kp1"], out[f"{tag}_kp2"], out[f"{tag}_i12"] = matches["kp1"], matches["kp2"], matches["i12"]
        out[f"{tag}_img_shape"] = np.asarray(matches["img_shape"], dtype=np.int64)
        for flags in ((1, 1, 1), (0, 0, 1), (1, 0, 0), (0, 1, 0)):
            res = sampson_case(ref, pose, matches, [bool(f) for f in flags])
            for k, v in res.items():
                out[f"{tag}_f{''.join(map(str, flags))}_{k}"] = v
    np.savez(os.path.join(OUT, "sampson.npz"), **out)


def golden_ggs(ref):
    """Full five-phase GGS call with iter_num=4 (phases 8/4/4/4/8) + an early-exit case."""
    out = {}
    cfg = syn.default_ggs_cfg()
    cfg["iter_num"] = 4
    for tag, seed, frames, per_pair in (("scene5", 31, 5, 96), ("scene8", 32, 8, 48)):
        m, gt, start = syn.scene_matches(frames, per_pair, seed=seed)
        res, text = quiet(ref.geometry_guided_sampling, torch.from_numpy(start)[None].clone(), 7, m, cfg)
        out[f"{tag}_pose"], out[f"{tag}_out"] = start, res[0].detach().numpy()
        out[f"{tag}_kp1"], out[f"{tag}_kp2"], out[f"{tag}_i12"] = m["kp1"], m["kp2"], m["i12"]
        out[f"{tag}_img_shape"] = np.asarray(m["img_shape"], dtype=np.int64)
        out[f"{tag}_log"] = np.array([float(l.split("sampson=")[1]) for l in text.splitlines() if l.startswith("t=")], dtype=np.float32)
        out[f"{tag}_drops"] = np.int64(sum("Drop" in l for l in text.splitlines()))
    # too few valid matches -> every phase stops before its first update
    m = syn.uniform_matches(5, 60, seed=33)
    pose = np.random.default_rng(33).normal(size=(5, 9)).astype(np.float32)
    res, text = quiet(ref.geometry_guided_sampling, torch.from_numpy(pose)[None].clone(), 3, m, cfg)
    out["drop_pose"], out["drop_out"] = pose, res[0].detach().numpy()
    out["drop_kp1"], out["drop_kp2"], out["drop_i12"] = m["kp1"], m["kp2"], m["i12"]
    out["drop_img_shape"] = np.asarray(m["img_shape"], dtype=np.int64)
    out["drop_drops"] = np.int64(sum("Drop" in l for l in text.splitlines()))
    out["iter_num"] = np.int64(cfg["iter_num"])
    np.savez(os.path.join(OUT, "ggs.npz"), **out)


def golden_loop(ref, sampler):
    """Free-running p_sample_loop, GGS off (N=5) and GGS on (N=5, iter_num=2), injected noise."""
    out = {}
    frames = 5
    z = syn.random_features(1, frames, seed=41)
    draws = syn.predraw_noise(1, frames, seed=41)
    out["z"], out["draws"] = z.numpy(), draws.numpy()

    def run(cond_fn, start_step):
        queue = [d.clone() for d in draws]
        orig_randn, orig_like = torch.randn, torch.randn_like
        torch.randn = lambda *a, **k: queue.pop(0)
        torch.randn_like = lambda *a, **k: queue.pop(0)
        try:
            (pose, trail), text = quiet(sampler.sample, [1, frames, 9], z, cond_fn, start_step)
        finally:
            torch.randn, torch.randn_like = orig_randn, orig_like
        return pose, trail, queue

    pose, trail, left = run(None, 0)
    out["off_pose"], out["off_trail"], out["off_unused_draws"] = pose.numpy(), trail.numpy(), np.int64(len(left))

    from functools import partial

    m, gt, start = syn.scene_matches(frames, 64, seed=42)
    cfg = syn.default_ggs_cfg()
    cfg["iter_num"] = 2
    cfg["min_matches"] = 0
    cond = partial(ref.geometry_guided_sampling, matches_dict=m, GGS_cfg=cfg)
    # the guided steps consume no draw: feed the reference only the draws it asks for
    pose, trail, left = run(cond, 10)
    out["on_pose"], out["on_trail"], out["on_unused_draws"] = pose.numpy(), trail.numpy(), np.int64(len(left))
    out["on_kp1"], out["on_kp2"], out["on_i12"] = m["kp1"], m["kp2"], m["i12"]
    out["on_img_shape"] = np.asarray(m["img_shape"], dtype=np.int64)
    out["on_iter_num"] = np.int64(2)
    np.savez(os.path.join(OUT, "loop.npz"), **out)


def synthetic_colmap_tables(frames=5, seed=51):
    """COLMAP-format inputs of colmap_keypoint_to_pytorch3d: float32 keypoints per (1-based) image, raw index matches per
    unordered pair (one pair without matches), crop boxes / scales as load_and_preprocess_images builds them."""
    rng = np.random.default_rng(seed)
    keypoints = {i + 1: rng.uniform(0, 1000, size=(int(rng.integers(40, 90)), 2)).astype(np.float32) for i in range(frames)}
    matches = {}
    for a in range(1, frames + 1):
        for b in range(a + 1, frames + 1):
            if (a, b) == (2, 4):
                matches[(a, b)] = None
                continue
            m = int(rng.integers(5, 60))
            matches[(a, b)] = np.stack([rng.integers(0, len(keypoints[a]), m), rng.integers(0, len(keypoints[b]), m)], 1).astype(np.int64)
    bboxes = np.stack([np.array([rng.integers(0, 300), rng.integers(0, 40), 0, 0], dtype=np.float32) for _ in range(frames)])
    bboxes[:, 2:] = bboxes[:, :2] + 1066
    image_info = {"size": (1066, 1066), "bboxes_xyxy": bboxes, "resized_scales": np.stack([224 / 1066] * frames)}
    return matches, keypoints, image_info


def golden_colmap():
    """Runs the reference's OWN colmap_keypoint_to_pytorch3d.  util/match_extraction.py imports hloc / pycolmap (absent) at
    module level, so the function's source is extracted from the file with `ast` and executed as is."""
    import ast

    path = os.path.join(os.environ.get("POSEDIFF_REFERENCE_ROOT", "/root/reference"), "pose_diffusion", "util", "match_extraction.py")
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "colmap_keypoint_to_pytorch3d")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    matches, keypoints, image_info = synthetic_colmap_tables()
    kp1, kp2, i12 = ns["colmap_keypoint_to_pytorch3d"](matches, {k: v.copy() for k, v in keypoints.items()}, image_info)
    np.savez(os.path.join(OUT, "colmap.npz"), kp1=kp1, kp2=kp2, i12=i12, numpy_version=np.__version__)


def main():
    torch.set_num_threads(1)  # bit-stable fixtures
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    state = syn.random_denoiser_state(WEIGHT_SEED, BIAS_STD)
    sampler = build_reference_sampler(ref, state)
    golden_schedule(ref)
    golden_denoiser(ref, sampler, state)
    golden_p_sample(ref, sampler)
    golden_sampson(ref)
    golden_ggs(ref)
    golden_loop(ref, sampler)
    golden_colmap()
    for name in sorted(os.listdir(OUT)):
        print(name, os.path.getsize(os.path.join(OUT, name)))


if __name__ == "__main__":
    main()
