"""Stage timing probe of the GGS kernel (debug helper, run on the GPU box)."""
import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import posediffusion_b200 as pdb
from posediffusion_b200 import synthetic as syn, _native
frames, per_pair = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (20, 2048)
dev = torch.device('cuda:0')
ctx = _native.Context.get(dev)
ctx.set_ggs_layout(sys.argv[3] if len(sys.argv) > 3 else 'paired')  # plain | paired (csrc/ggs_layout.cuh)
m = syn.uniform_matches(frames, per_pair, seed=0)
pm = ctx.pack_matches(m)
_, _, start = syn.scene_matches(frames, 4, seed=1)
cfg = syn.default_ggs_cfg(); cfg['min_matches'] = 0
pose = torch.from_numpy(start)[None].to(dev).clone()
for _ in range(2): ctx.ggs([pm], pose.clone(), cfg, want_stats=False)
torch.cuda.synchronize()
ctx.ggs_clocks(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ctx.ggs([pm], pose.clone(), cfg, want_stats=False); e1.record(); torch.cuda.synchronize()
clk = ctx.ggs_clocks(True, read=True)
used = clk[clk[:, 5] > 0]
it = used[0, 5]
print(f"launch {e0.elapsed_time(e1):.3f} ms, {it} iterations, {len(used)} CTAs, {e0.elapsed_time(e1)*1e3/it:.2f} us/iter")
names = ['stage3 norms', 'stage1', 'stage2b', 'exchange', 'stage2a', 'iters', 'next stage0', 'stage3 update']
for k, n in enumerate(names):
    if k == 5: continue
    per = used[:, k] / it
    print(f"{n:10s} cycles/iter: mean {per.mean():8.0f}  min {per.min():8.0f}  max {per.max():8.0f}")
print("total cycles/iter (cta0):", (used[0, :5].sum() + used[0, 6] + used[0, 7]) / it)
