"""Warm per-launch time of the tensor-core linear layer at the denoiser's shapes (GPU box only): 200 identical launches
captured into one CUDA graph (serialised on one stream), replayed 20 times, CUDA events around the replays.
    python tools/tc_gemm_probe.py [tokens=160]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from posediffusion_b200 import _native

S = int(sys.argv[1]) if len(sys.argv) > 1 else 160
ctx = _native.Context.get("cuda:0")
shapes = [("qkv", 1536, 512, False), ("out-proj (in place)", 512, 512, True), ("ff1", 1024, 512, False), ("ff2 (in place)", 512, 1024, True),
          ("last0", 128, 512, False)]
REP = 200
for name, O, K, in_place in shapes:
    x = torch.randn(S, K, device="cuda")
    w = torch.randn(O, K, device="cuda") * 0.02
    b = torch.randn(O, device="cuda")
    h = torch.zeros(S, O, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ctx.tc_linear(x, w, b, h if in_place else None, in_place=in_place)  # warm-up: attributes, tensor maps
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(REP):
                ctx.tc_linear(x, w, b, h if in_place else None, in_place=in_place)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (20 * REP)
    flops = 2.0 * S * O * K
    print(f"{name:22s} S {S:4d} O {O:4d} K {K:4d}: {us:6.2f} us per launch  ({flops / us * 1e-6:7.2f} TFLOP/s, weights {O * K * 4 / us * 1e-3:7.1f} GB/s)")
