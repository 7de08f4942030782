"""CPU oracle ("port") of PoseDiffusion's sampling hot path -- TEST INFRASTRUCTURE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  Nothing under `posediffusion_b200/` does; the product path
raises when its CUDA library is missing rather than falling back to this code.

It restates, in plain PyTorch-CPU fp32 with the same operator sequence as the reference
(so that it is also a fair CPU wall-clock baseline), the algorithm of
  * GaussianDiffusion schedule / p_sample / p_sample_loop  (models/gaussian_diffuser.py:120-300)
  * Denoiser.forward + embeddings                         (models/denoiser.py:53-76, util/embedding.py:13-50)
  * geometry_guided_sampling / GGS_optimize / compute_sampson_distance
                                                          (util/geometry_guided_sampling.py:14-172)
  * get_fundamental_matrices / pose_encoding_to_camera    (util/get_fundamental_matrix.py:14-51,
                                                           util/camera_transform.py:64-105)
plus the pytorch3d helpers those call (quaternion_to_matrix, opencv_from_cameras_projection,
hat, HarmonicEmbedding; pytorch3d is a third-party dependency that is absent from
/root/reference and UNPINNED in install.sh:24 -- we restate the 0.7.x semantics).

Pinning: the reference ships no tests or golden vectors for this path.
The oracle is pinned instead against the reference's OWN modules imported from
/root/reference in the build container (`oracle/make_golden.py` -> `tests/golden/*.npz`,
`tests/test_oracle_golden.py`).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

LOG_FL_BIAS = 1.8  # camera_transform.py:67
FL_MIN, FL_MAX = 0.1, 20.0  # camera_transform.py:68-69

SCHEDULE_KEYS = (
    "betas",
    "alphas_cumprod",
    "alphas_cumprod_prev",
    "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod",
    "log_one_minus_alphas_cumprod",
    "sqrt_recip_alphas_cumprod",
    "sqrt_recipm1_alphas_cumprod",
    "posterior_variance",
    "posterior_log_variance_clipped",
    "posterior_mean_coef1",
    "posterior_mean_coef2",
    "p2_loss_weight",
)


# ----------------------------------------------------------------------------------------
# DDPM schedule  (gaussian_diffuser.py:136-187, "custom" = linspace(beta_1, beta_T) in fp64)
# ----------------------------------------------------------------------------------------
def diffusion_schedule(timesteps: int = 100, beta_1: float = 1e-4, beta_T: float = 0.1) -> Dict[str, torch.Tensor]:
    beta = torch.linspace(beta_1, beta_T, timesteps, dtype=torch.float64)
    alpha = 1.0 - beta
    abar = torch.cumprod(alpha, dim=0)
    abar_prev = torch.cat([torch.ones(1, dtype=torch.float64), abar[:-1]])
    post_var = beta * (1.0 - abar_prev) / (1.0 - abar)
    table = {
        "betas": beta,
        "alphas_cumprod": abar,
        "alphas_cumprod_prev": abar_prev,
        "sqrt_alphas_cumprod": abar.sqrt(),
        "sqrt_one_minus_alphas_cumprod": (1.0 - abar).sqrt(),
        "log_one_minus_alphas_cumprod": (1.0 - abar).log(),
        "sqrt_recip_alphas_cumprod": (1.0 / abar).sqrt(),
        "sqrt_recipm1_alphas_cumprod": (1.0 / abar - 1).sqrt(),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": post_var.clamp(min=1e-20).log(),
        "posterior_mean_coef1": beta * abar_prev.sqrt() / (1.0 - abar),
        "posterior_mean_coef2": (1.0 - abar_prev) * alpha.sqrt() / (1.0 - abar),
        "p2_loss_weight": (1 + abar / (1 - abar)) ** -0.0,
    }
    return {k: v.to(torch.float32) for k, v in table.items()}


# ----------------------------------------------------------------------------------------
# Embeddings  (embedding.py:24-37, :40-50 + pytorch3d HarmonicEmbedding)
# ----------------------------------------------------------------------------------------
def timestep_features(t: torch.Tensor, dim: int = 256, max_period: float = 10000.0) -> torch.Tensor:
    """[B] int -> [B, dim] = [cos(t f_k) | sin(t f_k)], f_k = max_period^(-k/half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    phase = t[:, None].float() * freqs[None]
    return torch.cat([phase.cos(), phase.sin()], dim=-1)


def harmonic_features(x: torch.Tensor, n_harmonic: int = 10) -> torch.Tensor:
    """[..., C] -> [..., C*(2n+1)]: sin block | cos block | x, channel-major / octave-minor."""
    octaves = 2.0 ** torch.arange(n_harmonic, dtype=torch.float32)
    arg = (x[..., None] * octaves).reshape(*x.shape[:-1], -1)
    return torch.cat([arg.sin(), arg.cos(), x], dim=-1)


# ----------------------------------------------------------------------------------------
# Denoiser  (denoiser.py:22-98).  Built from the same torch.nn library modules the
# reference composes (nn.TransformerEncoder pre-norm, ReLU FFN), so state_dicts interchange.
# ----------------------------------------------------------------------------------------
class _TimeEmbed(nn.Module):
    def __init__(self):
        super().__init__()
        self.linear = nn.Sequential(nn.Linear(256, 128), nn.SiLU(), nn.Linear(128, 128))

    def forward(self, t):
        return self.linear(timestep_features(t))


class OracleDenoiser(nn.Module):
    def __init__(self, d_model=512, nhead=4, dim_feedforward=1024, num_layers=8, target_dim=9, z_dim=384, hidden=128):
        super().__init__()
        self.target_dim = target_dim
        self.time_embed = _TimeEmbed()
        in_dim = 128 + target_dim * 21 + z_dim + 1
        self._first = nn.Linear(in_dim, d_model)
        layer = nn.TransformerEncoderLayer(
            d_model=d_model, nhead=nhead, dim_feedforward=dim_feedforward, dropout=0.1, batch_first=True, norm_first=True
        )
        self._trunk = nn.TransformerEncoder(layer, num_layers)
        self._last = nn.Sequential(nn.Linear(d_model, hidden), nn.LayerNorm(hidden), nn.ReLU(), nn.Linear(hidden, target_dim))

    def forward(self, x: torch.Tensor, t: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        batch, frames, _ = x.shape
        t_tok = self.time_embed(t)[:, None, :].expand(-1, frames, -1)
        pivot = torch.zeros(batch, frames, 1, dtype=z.dtype)
        pivot[:, 0] = 1.0  # one-hot of the first (pivot) camera, denoiser.py:62-66
        tokens = torch.cat([harmonic_features(x), t_tok, z, pivot], dim=-1)  # column order denoiser.py:68
        return self._last(self._trunk(self._first(tokens)))


def build_denoiser(state: Dict[str, torch.Tensor]) -> OracleDenoiser:
    net = OracleDenoiser()
    net.load_state_dict({k: v.clone() for k, v in state.items()}, strict=True)
    return net.eval()


# ----------------------------------------------------------------------------------------
# Sampler  (gaussian_diffuser.py:190-300) with injected Gaussian draws
# ----------------------------------------------------------------------------------------
def posterior_mean(sched: Dict[str, torch.Tensor], x: torch.Tensor, eps: torch.Tensor, t: int):
    """x0 = a_t x - b_t eps ; mean = c1_t x0 + c2_t x   (:190-194, :201-209)."""
    x0 = sched["sqrt_recip_alphas_cumprod"][t] * x - sched["sqrt_recipm1_alphas_cumprod"][t] * eps
    mean = sched["posterior_mean_coef1"][t] * x0 + sched["posterior_mean_coef2"][t] * x
    return mean, x0


@torch.no_grad()
def p_sample(net, sched, x, t: int, z, noise, cond_fn=None, cond_start_step: int = 0):
    """One ancestral step (:249-282).  `noise` is the pre-drawn N(0,1) tensor for this step."""
    steps = torch.full((x.shape[0],), t, dtype=torch.long)
    eps = net(x, steps, z)
    mean, x0 = posterior_mean(sched, x, eps, t)
    if cond_fn is not None and t < cond_start_step:
        mean = cond_fn(mean, t)
        return mean, x0  # guided steps add no noise (:270-276)
    if t == 0:
        return mean, x0
    sigma = (0.5 * sched["posterior_log_variance_clipped"][t]).exp()
    return mean + sigma * noise, x0


@torch.no_grad()
def p_sample_loop(net, sched, z, draws: torch.Tensor, cond_fn=None, cond_start_step: int = 0, timesteps: int = 100):
    """draws[0] = x_T, draws[1+k] = noise for loop iteration k (t = timesteps-1-k).
    Returns (pose [B,N,9], trajectory [T+1,B,N,9])  (:285-300)."""
    pose = draws[0].clone()
    trail = [pose.clone()]
    for k, t in enumerate(reversed(range(timesteps))):
        pose, _ = p_sample(net, sched, pose, t, z, draws[1 + k], cond_fn, cond_start_step)
        trail.append(pose.clone())
    return pose, torch.stack(trail)


# ----------------------------------------------------------------------------------------
# Geometry  (camera_transform.py:64-105, get_fundamental_matrix.py:14-51 + pytorch3d)
# ----------------------------------------------------------------------------------------
def quat_to_rotation(q: torch.Tensor) -> torch.Tensor:
    w, x, y, z = q.unbind(-1)
    s2 = 2.0 / (q * q).sum(-1)
    entries = (
        1 - s2 * (y * y + z * z), s2 * (x * y - z * w), s2 * (x * z + y * w),
        s2 * (x * y + z * w), 1 - s2 * (x * x + z * z), s2 * (y * z - x * w),
        s2 * (x * z - y * w), s2 * (y * z + x * w), 1 - s2 * (x * x + y * y),
    )
    return torch.stack(entries, -1).reshape(q.shape[:-1] + (3, 3))


def skew(v: torch.Tensor) -> torch.Tensor:
    x, y, z = v.unbind(-1)
    o = torch.zeros_like(x)
    return torch.stack((o, -z, y, z, o, -x, -y, x, o), dim=-1).reshape(v.shape[:-1] + (3, 3))


def decode_pose(pose: torch.Tensor):
    """[.., N, 9] -> (R [n,3,3], T [n,3], focal [n,2])  (camera_transform.py:85-97)."""
    flat = pose.reshape(-1, pose.shape[-1])
    focal = torch.clamp((flat[:, 7:9] + LOG_FL_BIAS).exp(), min=FL_MIN, max=FL_MAX)
    return quat_to_rotation(flat[:, 3:7]), flat[:, :3], focal


def fundamental_all_pairs(R, T, focal, height: int, width: int, idx1, idx2):
    """F for the listed (idx1, idx2) camera pairs, same op order as the reference:
    NDC->OpenCV, R12, t12, E = R12 hat(-R12^T t12), F = K2^-T E K1^-1 (batched inverse)."""
    n = R.shape[0]
    flip = torch.tensor([-1.0, -1.0, 1.0])
    R_cv = (R * flip[None, None, :]).permute(0, 2, 1)
    t_cv = T * flip[None, :]
    scale = min(height, width) / 2.0
    K = torch.zeros(n, 3, 3, dtype=R.dtype)
    K[:, 0, 0] = focal[:, 0] * scale
    K[:, 1, 1] = focal[:, 1] * scale
    K[:, 0, 2] = width / 2.0
    K[:, 1, 2] = height / 2.0
    K[:, 2, 2] = 1.0
    R1, t1, K1 = R_cv[idx1], t_cv[idx1], K[idx1]
    R2, t2, K2 = R_cv[idx2], t_cv[idx2], K[idx2]
    R12 = R2.matmul(R1.permute(0, 2, 1))
    t12 = t2 - R12.matmul(t1[..., None])[..., 0]
    e_t = -R12.permute(0, 2, 1).matmul(t12[..., None])[..., 0]
    E = R12.matmul(skew(e_t))
    return K2.inverse().permute(0, 2, 1).matmul(E).matmul(K1.inverse())


def prepare_matches(matches: Dict) -> Dict:
    """geometry_guided_sampling.py:16-45: homogeneous points, pair index a*N+b, N x N meshgrid."""
    frames, _, height, width = matches["img_shape"]
    kp1 = torch.from_numpy(np.ascontiguousarray(matches["kp1"]))
    kp2 = torch.from_numpy(np.ascontiguousarray(matches["kp2"]))
    i12 = torch.from_numpy(np.ascontiguousarray(matches["i12"]))
    grid = torch.arange(frames)
    return {
        "x1": F.pad(kp1, [0, 1], value=1),
        "x2": F.pad(kp2, [0, 1], value=1),
        "pair": (i12[:, 0] * frames + i12[:, 1]).long(),
        "idx1": grid.repeat_interleave(frames),
        "idx2": grid.repeat(frames),
        "h": height,
        "w": width,
        "frames": frames,
    }


def sampson_terms(pose: torch.Tensor, prep: Dict, update_R=True, update_T=True, update_FL=True, sampson_max=10):
    """compute_sampson_distance (:129-172): returns (valid errors, clamped mean for logging)."""
    R, T, focal = decode_pose(pose)
    focal = focal.mean(dim=0).repeat(len(focal), 1)  # shared focal length (:142)
    if not update_R:
        R = R.detach()
    if not update_T:
        T = T.detach()
    if not update_FL:
        focal = focal.detach()
    Fm = fundamental_all_pairs(R, T, focal, prep["h"], prep["w"], prep["idx1"], prep["idx2"]).permute(0, 2, 1)
    x1 = prep["x1"].float()
    x2 = prep["x2"].float()
    Fpm = Fm[prep["pair"]]
    left = torch.bmm(x1[:, None], Fpm)
    right = torch.bmm(Fpm, x2[..., None])
    denom = left[:, :, 0].square() + left[:, :, 1].square() + right[:, 0, :].square() + right[:, 1, :].square()
    numer = torch.bmm(left, x2[..., None]).square()
    err = numer[:, 0] / denom
    logged = err.detach().clone().clamp(max=sampson_max).mean()
    return err[err < sampson_max], logged


def ggs_phase(
    pose: torch.Tensor,
    prep: Dict,
    update_R=True,
    update_T=True,
    update_FL=True,
    alpha=1e-4,
    learning_rate=1e-2,
    iter_num=100,
    sampson_max=10,
    min_matches=10,
    momentum=0.9,
    log: Optional[List] = None,
    **_unused,
) -> torch.Tensor:
    """GGS_optimize (:67-126): SGD-momentum on the posterior mean with norm-relative clipping.
    The SGD / clip_grad_norm_ arithmetic is written out (torch.optim.SGD, momentum 0.9, dampening 0;
    clip coefficient min(1, max_norm / (|g| + 1e-6)))."""
    if update_R and update_T and update_FL:
        iter_num = iter_num * 2
    frames = pose.shape[1]
    pose = pose.detach().clone()
    velocity = None
    logged = torch.tensor(float("nan"))
    done = 0
    dropped = False
    for _ in range(iter_num):
        with torch.enable_grad():
            leaf = pose.clone().requires_grad_(True)
            valid, logged = sampson_terms(leaf, prep, update_R, update_T, update_FL, sampson_max)
            if min_matches > 0 and len(valid) / frames < min_matches:
                dropped = True
                break
            valid.mean().backward()
        grad = leaf.grad
        mask = grad.abs() > 0
        max_norm = alpha * (pose * mask).norm() / learning_rate
        coef = torch.clamp(max_norm / (grad.norm() + 1e-6), max=1.0)
        grad = grad * coef
        velocity = grad.clone() if velocity is None else velocity * momentum + grad
        pose = pose - learning_rate * velocity
        done += 1
    if log is not None:
        log.append({"sampson": float(logged), "iters": done, "dropped": dropped})
    return pose


def geometry_guided_sampling(pose: torch.Tensor, t: int, matches_dict: Dict, GGS_cfg: Dict, log: Optional[List] = None):
    """Five phases: all x2, focal only, rotation only, translation only, all x2 (:47-64)."""
    prep = prepare_matches(matches_dict)
    cfg = {k: v for k, v in GGS_cfg.items() if k not in ("enable", "start_step", "pose_encoding_type")}
    for flags in (
        dict(),
        dict(update_T=False, update_R=False, update_FL=True),
        dict(update_T=False, update_R=True, update_FL=False),
        dict(update_T=True, update_R=False, update_FL=False),
        dict(),
    ):
        pose = ggs_phase(pose, prep, log=log, **flags, **cfg)
    return pose
