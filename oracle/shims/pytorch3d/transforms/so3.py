import torch


def hat(v):
    """[n,3] -> [n,3,3] skew matrices: hat(v) @ u = v x u."""
    x, y, z = v.unbind(1)
    zero = torch.zeros_like(x)
    return torch.stack((zero, -z, y, z, zero, -x, -y, x, zero), dim=1).reshape(-1, 3, 3)
