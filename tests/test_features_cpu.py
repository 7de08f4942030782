"""CPU checks of the image-feature row: the oracle against the reference-generated golden vectors and an
independent ViT implementation, the host-only position-table helper, and the host mirror's checkpoint layout."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.dino_vit import DinoViTSmall16, multiscale_features, randomize
from oracle.make_golden_features import CASES, VIT_SEED, images_for
from posediffusion_b200 import _native


@pytest.fixture(scope="module")
def net():
    return randomize(DinoViTSmall16(), VIT_SEED).eval()


def test_seeded_weights_are_the_ones_the_fixture_was_made_with(net, golden):
    g = golden("features.npz")
    checksum = sum(float(v.double().abs().sum()) for v in net.state_dict().values())
    assert math.isclose(checksum, float(g["weight_checksum"]), rel_tol=1e-12)


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_wrapper_matches_reference_wrapper_golden(net, golden, case):
    """oracle.multiscale_features == the reference's MultiScaleImageFeatureExtractor (fixture made by its unmodified code).
    Same backbone module on both sides, so only the wrapper arithmetic is compared: 1e-5 (thread-count dependent sums)."""
    n, h, w, sf, seed = CASES[case]
    with torch.no_grad():
        z = multiscale_features(net, images_for(n, h, w, seed), sf).numpy()
    np.testing.assert_allclose(z, golden("features.npz")[case], rtol=0, atol=1e-5)


def test_restated_backbone_matches_independent_vit_implementation(net):
    """The block arithmetic of oracle/dino_vit.py against transformers.ViTModel (same architecture, independent code):
    class-token feature of a 224^2 batch within 2e-5."""
    transformers = pytest.importorskip("transformers")
    cfg = transformers.ViTConfig(hidden_size=384, num_hidden_layers=12, num_attention_heads=6, intermediate_size=1536,
                                 hidden_act="gelu", layer_norm_eps=1e-6, image_size=224, patch_size=16, qkv_bias=True,
                                 hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    hf = transformers.ViTModel(cfg, add_pooling_layer=False).eval()
    sd = net.state_dict()
    m = {"embeddings.cls_token": sd["cls_token"], "embeddings.position_embeddings": sd["pos_embed"],
         "embeddings.patch_embeddings.projection.weight": sd["patch_embed.proj.weight"],
         "embeddings.patch_embeddings.projection.bias": sd["patch_embed.proj.bias"],
         "layernorm.weight": sd["norm.weight"], "layernorm.bias": sd["norm.bias"]}
    for i in range(12):
        p, q = f"blocks.{i}.", f"encoder.layer.{i}."
        for j, name in enumerate(["query", "key", "value"]):
            m[q + f"attention.attention.{name}.weight"] = sd[p + "attn.qkv.weight"][j * 384:(j + 1) * 384]
            m[q + f"attention.attention.{name}.bias"] = sd[p + "attn.qkv.bias"][j * 384:(j + 1) * 384]
        m[q + "attention.output.dense.weight"], m[q + "attention.output.dense.bias"] = sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"]
        m[q + "layernorm_before.weight"], m[q + "layernorm_before.bias"] = sd[p + "norm1.weight"], sd[p + "norm1.bias"]
        m[q + "layernorm_after.weight"], m[q + "layernorm_after.bias"] = sd[p + "norm2.weight"], sd[p + "norm2.bias"]
        m[q + "intermediate.dense.weight"], m[q + "intermediate.dense.bias"] = sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]
        m[q + "output.dense.weight"], m[q + "output.dense.bias"] = sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"]
    hf.load_state_dict(m, strict=True)
    x = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        a = net(x)
        b = hf(pixel_values=x).last_hidden_state[:, 0]
    assert (a - b).abs().max().item() < 2e-5


@pytest.mark.parametrize("gh,gw", [(7, 7), (4, 4), (12, 14), (6, 7), (14, 14), (9, 5)])
def test_position_table_helper_matches_torch_bicubic(gh, gw):
    """pdb_vit_pos_table (host C++) against the hub model's interpolate_pos_encoding evaluated with torch: 5e-6."""
    pos = torch.randn(1, 197, 384, generator=torch.Generator().manual_seed(gh * 31 + gw))
    out = _native.vit_pos_table(pos.numpy(), gh, gw)
    if (gh, gw) == (14, 14):
        ref = pos[0]
    else:
        grid = pos[:, 1:].reshape(1, 14, 14, 384).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, scale_factor=((gh + 0.1) / 14, (gw + 0.1) / 14), mode="bicubic")
        assert tuple(grid.shape[-2:]) == (gh, gw)
        ref = torch.cat((pos[:, :1], grid.permute(0, 2, 3, 1).reshape(1, -1, 384)), dim=1)[0]
    np.testing.assert_allclose(out, ref.numpy(), rtol=0, atol=5e-6)


def test_host_mirror_has_the_checkpoint_layout_and_no_cpu_path(net):
    import posediffusion_b200 as pdb

    ext = pdb.MultiScaleImageFeatureExtractor(modelname="dino_vits16", freeze=True, scale_factors=[1, 1 / 2, 1 / 3])
    assert ext.get_output_dim() == 384
    assert all(not p.requires_grad for p in ext.parameters())
    ext._net.load_state_dict(net.state_dict(), strict=True)  # hub checkpoint names / shapes
    assert [tuple(v.shape) for v in ext._net.state_dict().values()] == [tuple(v.shape) for v in net.state_dict().values()]
    with pytest.raises(_native.NativeError):
        ext(torch.rand(1, 3, 224, 224))  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        pdb.MultiScaleImageFeatureExtractor(modelname="something_else")
    empty = pdb.MultiScaleImageFeatureExtractor(scale_factors=[])
    with pytest.raises(ValueError):
        empty(torch.rand(1, 3, 224, 224))
