"""Per-stage error of the CUDA image backbone against the CPU oracle + a timing line (development probe)."""
import sys, time
import numpy as np
import torch

sys.path.insert(0, ".")
from oracle.dino_vit import DinoViTSmall16, multiscale_features, randomize
from oracle.make_golden_features import CASES, VIT_SEED, images_for
from posediffusion_b200 import _native

net = randomize(DinoViTSmall16(), VIT_SEED).eval()
ctx = _native.Context.get("cuda:0")
ctx.load_vit([v.cuda() for v in net.state_dict().values()])
n, h, w, sf, seed = CASES["default"]
if "--sanitize" in sys.argv:  # small shapes for compute-sanitizer: backbone on 2 frames + the post-loop kernels
    x = torch.rand(2, 3, 224, 224, device="cuda")
    z = ctx.extract_features(x, sf)
    z2 = ctx.extract_features(torch.rand(1, 3, 192, 224, device="cuda"), sf)
    R, T, F = ctx.pose_to_camera(torch.randn(2, 5, 9, device="cuda"))
    r, t = ctx.rel_pose_error(R, T, R.clone(), T.clone() + 0.1, 2)
    torch.cuda.synchronize()
    print("sanitize run ok", float(z.abs().mean()), float(z2.abs().mean()), float(r.mean()), float(t.mean()))
    sys.exit(0)
if "--profile" in sys.argv:  # two calls at 20 frames for an ncu launch list
    x = torch.rand(20, 3, 224, 224, device="cuda")
    ctx.extract_features(x, sf)
    ctx.extract_features(x, sf)
    torch.cuda.synchronize()
    sys.exit(0)
img = images_for(n, h, w, seed)
with torch.no_grad():
    zref, stages = multiscale_features(net, img, sf, return_stages=True)
for k in range(13):
    z, dbg = ctx.extract_features(img.cuda(), sf, debug_stage=k)
    ref = torch.cat([st[k].reshape(-1, 384) for st in stages], 0)
    err = (dbg.cpu() - ref).abs()
    per_scale = [err[a:b].max().item() for a, b in [(0, n * 197), (n * 197, n * 247), (n * 247, n * 264)]]
    print(f"stage {k:2d} max|err| {err.max().item():.3e} per-scale {per_scale} ref max {ref.abs().max().item():.2f} cls-row err {err[0].max().item():.2e}")
print("z err", (z.cpu() - zref).abs().max().item())
for nimg in (5, 20, 80):
    x = torch.rand(nimg, 3, 224, 224, device="cuda")
    for _ in range(3):
        ctx.extract_features(x, sf)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ctx.extract_features(x, sf)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    flops = nimg * 264 * 2 * (768 * 384 + 12 * (384 * 1152 + 384 * 384 + 2 * 384 * 1536))
    print(f"n={nimg}: {ms:.3f} ms/call, {nimg / ms * 1e3:.0f} images/s, GEMM {flops / ms / 1e9:.1f} TFLOP/s")
