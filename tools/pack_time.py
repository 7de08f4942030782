import sys, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from posediffusion_b200 import synthetic as syn, _native
ctx=_native.Context.get('cuda:0')
m=syn.uniform_matches(20,2048,seed=0)
for i in range(4):
    t0=time.perf_counter(); pm=ctx.pack_matches(m); torch.cuda.synchronize(); t1=time.perf_counter()
    print('pack ms', (t1-t0)*1e3)
    del pm
import ctypes
kp1=np.ascontiguousarray(m['kp1']); 
t0=time.perf_counter(); a=np.ascontiguousarray(m['kp1'],dtype=np.float64).reshape(-1,2); b=np.ascontiguousarray(m['i12'],dtype=np.int64).reshape(-1,2); t1=time.perf_counter(); print('numpy prep ms',(t1-t0)*1e3)
