"""Generate tests/golden/features.npz with the REFERENCE's own MultiScaleImageFeatureExtractor.

TEST INFRASTRUCTURE.  Run in the build container only:  python -m oracle.make_golden_features
`models/image_feature_extractor.py` is imported unmodified from /root/reference; its `torch.hub.load` call (the backbone is a
third-party download, absent here) is redirected to the restated network of oracle/dino_vit.py with seeded parameters.
Normalisation, the bilinear pyramid and the averaging over scales are therefore the reference's code.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.dino_vit import DinoViTSmall16, randomize  # noqa: E402
from oracle.ref_loader import REFERENCE_ROOT  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "features.npz")
VIT_SEED = 11
CASES = {  # name: (n_images, H, W, scale_factors, image seed)
    "default": (2, 224, 224, [1, 1 / 2, 1 / 3], 101),
    "single_scale": (1, 224, 224, [1], 102),
    "non_square": (2, 192, 224, [1, 1 / 2, 1 / 3], 103),
}


def images_for(n, h, w, seed):
    return torch.rand((n, 3, h, w), generator=torch.Generator().manual_seed(seed))


def reference_extractor(scale_factors):
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "pose_diffusion", "models"))
    import image_feature_extractor as ref_mod  # the reference file itself

    real = torch.hub.load
    torch.hub.load = lambda repo, name, *a, **k: randomize(DinoViTSmall16(), VIT_SEED)
    try:
        ext = ref_mod.MultiScaleImageFeatureExtractor(modelname="dino_vits16", scale_factors=list(scale_factors))
    finally:
        torch.hub.load = real
    return ext.eval()


def main():
    torch.set_num_threads(1)
    out = {}
    for name, (n, h, w, sf, seed) in CASES.items():
        ext = reference_extractor(sf)
        with torch.no_grad():
            out[name] = ext(images_for(n, h, w, seed)).numpy()
        print(name, out[name].shape, float(np.abs(out[name]).mean()))
    net = randomize(DinoViTSmall16(), VIT_SEED)
    out["weight_checksum"] = np.float64(sum(float(v.double().abs().sum()) for v in net.state_dict().values()))
    np.savez(OUT, **out)
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
