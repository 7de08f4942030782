"""Stage timing probe of the fp32 persistent denoiser kernel (debug helper, run on the GPU box).

    python tools/den_stage_probe.py [frames=20] [batch=1]
"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import posediffusion_b200 as pdb
from posediffusion_b200 import synthetic as syn

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device('cuda:0')
den = pdb.Denoiser(TRANSFORMER=dict(d_model=512, nhead=4, dim_feedforward=1024, num_encoder_layers=8, dropout=0.1, batch_first=True, norm_first=True))
den.load_state_dict(syn.random_denoiser_state(0), strict=True)
den = den.to(dev)
ctx = den.native_context()
ctx.set_denoiser_engine('fp32')
z = syn.random_features(batch, frames, 0).to(dev)
draws = syn.predraw_noise(batch, frames, seed=0).to(dev)
for _ in range(2):
    ctx.sample_loop(z, draws, None, None, 0, want_trail=False, want_stats=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ctx.sample_loop(z, draws, None, None, 0, want_trail=False, want_stats=False); e1.record(); torch.cuda.synchronize()
print(f"unprobed loop {e0.elapsed_time(e1):.3f} ms = {e0.elapsed_time(e1) * 10:.1f} us per step")
ctx.ggs_clocks(2)
e0.record(); ctx.sample_loop(z, draws, None, None, 0, want_trail=False, want_stats=False); e1.record(); torch.cuda.synchronize()
clk = ctx.ggs_clocks(2, read=True)
ctx.ggs_clocks(False)
used = clk[clk[:, 5] > 0]
steps = used[0, 5]
print(f"probed loop {e0.elapsed_time(e1):.3f} ms, {steps} steps, {len(used)} CTAs")
for k, n in enumerate(['barrier (43 per step)', 'tile load + LayerNorm', 'linear item (FMA + reduce + store)', 'attention', 'tail']):
    per = used[:, k] / steps
    print(f"{n:36s} cycles/step: mean {per.mean():9.0f}  min {per.min():9.0f}  max {per.max():9.0f}")
print("sum (cta0) cycles/step:", used[0, :5].sum() / steps)
