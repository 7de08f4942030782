import torch


class CamerasBase:
    pass


class PerspectiveCameras(CamerasBase):
    """Plain container: R [n,3,3], T [n,3], focal_length [n,2], principal_point 0 (NDC)."""

    def __init__(self, focal_length=1.0, principal_point=None, R=None, T=None, device="cpu", in_ndc=True, **_):
        self.R = R
        self.T = T
        self.focal_length = focal_length
        count = R.shape[0]
        if principal_point is None:
            principal_point = torch.zeros(count, 2, dtype=R.dtype, device=R.device)
        self.principal_point = principal_point
        self.device = R.device

    def __len__(self):
        return self.R.shape[0]
