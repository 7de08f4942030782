"""Minimal stand-in for `hydra` (absent from this image, no network).

TEST INFRASTRUCTURE ONLY: lets `oracle/ref_loader.py` import the reference's
own modules from /root/reference unchanged.  Only `hydra.utils.instantiate`
(reference call sites: denoiser.py:48, pose_diffusion_model.py:57-60) is needed.
"""
