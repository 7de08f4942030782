"""Tensor-core engine below 128 tokens (swap-AB tiles) against the exact-fp32 persistent kernel: time per diffusion step.

    python tools/tc_small_probe.py            # prints one JSON line; run once more with PDB_TC_SWAP=1 for the swap-AB tiles
"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import posediffusion_b200 as pdb
from posediffusion_b200 import synthetic as syn

dev = torch.device('cuda:0')
den = pdb.Denoiser(TRANSFORMER=dict(d_model=512, nhead=4, dim_feedforward=1024, num_encoder_layers=8, dropout=0.1, batch_first=True, norm_first=True))
den.load_state_dict(syn.random_denoiser_state(0), strict=True)
den = den.to(dev)
ctx = den.native_context()
out = {"tc_swap": os.environ.get("PDB_TC_SWAP", "0")}
for batch, frames in ((1, 5), (1, 20), (1, 80), (4, 20)):
    z = syn.random_features(batch, frames, 0).to(dev)
    draws = syn.predraw_noise(batch, frames, seed=0).to(dev)
    for engine in ("fp32", "tf32"):
        ctx.set_denoiser_engine(engine)
        for _ in range(2):
            ctx.sample_loop(z, draws, None, None, 0, want_trail=False, want_stats=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ctx.sample_loop(z, draws, None, None, 0, want_trail=False, want_stats=False)
        e1.record()
        torch.cuda.synchronize()
        out[f"{batch}x{frames}_{engine}_us_per_step"] = round(e0.elapsed_time(e1) / 3 * 10, 1)
ctx.set_denoiser_engine("auto")
print(json.dumps(out))
