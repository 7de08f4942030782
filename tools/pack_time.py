"""Timing of the match-ingestion / host-buffer entry points (debug helper, run on the GPU box)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import posediffusion_b200 as pdb
from posediffusion_b200 import synthetic as syn, _native

dev = torch.device('cuda:0')
den = pdb.Denoiser(TRANSFORMER=dict(d_model=512, nhead=4, dim_feedforward=1024, num_encoder_layers=8, dropout=0.1, batch_first=True, norm_first=True))
den.load_state_dict(syn.random_denoiser_state(0), strict=True)
ctx = den.to(dev).native_context()
m = syn.uniform_matches(20, 2048, seed=0)
for i in range(4):
    t0 = time.perf_counter(); pm = ctx.pack_matches(m); torch.cuda.synchronize(); t1 = time.perf_counter()
    print('pack ms', round((t1 - t0) * 1e3, 2))
    t0 = time.perf_counter(); del pm; torch.cuda.synchronize(); print('free ms', round((time.perf_counter() - t0) * 1e3, 2))
cfg = syn.default_ggs_cfg(); cfg['verbose'] = False
z = syn.random_features(1, 20, 0).pin_memory(); draws = syn.predraw_noise(1, 20, seed=0).pin_memory(); out = torch.empty(1, 20, 9).pin_memory()
pm = ctx.pack_matches(m)
for i in range(3):
    t0 = time.perf_counter(); ctx.sample_loop_host(z.numpy(), draws.numpy(), [pm], cfg, 10, out.numpy()); t1 = time.perf_counter()
    print('sample_loop_host ms', round((t1 - t0) * 1e3, 2))
zd, dd = z.to(dev), draws.to(dev)
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); ctx.sample_loop(zd, dd, [pm], cfg, 10, want_trail=False, want_stats=False); torch.cuda.synchronize(); t1 = time.perf_counter()
    print('sample_loop (device) ms', round((t1 - t0) * 1e3, 2))
