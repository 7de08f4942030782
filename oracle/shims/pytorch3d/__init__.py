"""Minimal stand-in for `pytorch3d` (absent from this image; the reference leaves its
version unpinned, install.sh:24).

TEST INFRASTRUCTURE ONLY.  Restates, from the published pytorch3d API semantics, the six
symbols the reference's sampling hot path calls:
quaternion_to_matrix, PerspectiveCameras, opencv_from_cameras_projection, hat,
HarmonicEmbedding (+ names that only need to exist at import time).
"""
__version__ = "0.7-shim"
