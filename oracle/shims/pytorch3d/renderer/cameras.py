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

    def get_world_to_view_transform(self):
        """pytorch3d row-vector convention: X_view = X_world @ R + T, i.e. the 4x4 matrix [[R, 0], [T, 1]]."""
        R, T = torch.as_tensor(self.R), torch.as_tensor(self.T)
        m = torch.zeros(R.shape[0], 4, 4, dtype=R.dtype, device=R.device)
        m[:, :3, :3] = R
        m[:, 3, :3] = T
        m[:, 3, 3] = 1.0
        return _Matrix(m)


class _Matrix:
    def __init__(self, m):
        self._m = m

    def get_matrix(self):
        return self._m
