"""tcgen05 + TMA linear layer (csrc/tc_linear.cuh) vs a plain PyTorch fp32 matmul of the same op.
Tolerance: TF32 products (10-bit mantissa), fp32 accumulate -> 2e-3 of max |y| (stated; K <= 1024)."""
import numpy as np
import pytest
import torch

from posediffusion_b200 import _native

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    assert torch.cuda.is_available()
    return _native.Context.get("cuda:0")


@pytest.mark.parametrize("S,O,K", [(128, 64, 32), (128, 512, 512), (160, 1536, 512), (20, 512, 1024), (300, 1024, 512), (1280, 512, 384)])
@pytest.mark.parametrize("epi", ["plain", "bias_res_relu"])
def test_tc_linear_matches_fp32(ctx, S, O, K, epi):
    g = torch.Generator(device="cuda").manual_seed(S + O + K)
    x = torch.randn(S, K, device="cuda", generator=g)
    w = torch.randn(O, K, device="cuda", generator=g) * 0.05
    bias = torch.randn(O, device="cuda", generator=g) if epi != "plain" else None
    res = torch.randn(S, O, device="cuda", generator=g) if epi != "plain" else None
    y = ctx.tc_linear(x, w, bias, res, relu=(epi != "plain"))
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = x.double() @ w.double().T
    if bias is not None:
        ref = torch.relu(ref + bias.double() + res.double())
    err = (y.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale, (err, scale)
    # and it is really TF32-accurate, not garbage that happens to be small: correlation with the fp64 result
    assert torch.corrcoef(torch.stack([y.double().flatten(), ref.flatten()]))[0, 1].item() > 0.99999
