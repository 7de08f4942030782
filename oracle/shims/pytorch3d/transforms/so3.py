import torch


def hat(v):
    """[n,3] -> [n,3,3] skew matrices: hat(v) @ u = v x u."""
    x, y, z = v.unbind(1)
    zero = torch.zeros_like(x)
    return torch.stack((zero, -z, y, z, zero, -x, -y, x, zero), dim=1).reshape(-1, 3, 3)


def acos_linear_extrapolation(x, bounds=(-1.0 + 1e-4, 1.0 - 1e-4)):
    """pytorch3d.transforms.math.acos_linear_extrapolation: acos inside the bounds, first-order Taylor extension outside
    (so that the derivative stays finite at +-1)."""
    lower, upper = bounds
    import math

    def _dacos(v):
        return -1.0 / math.sqrt(1.0 - v * v)

    out = torch.empty_like(x)
    hi, lo = x > upper, x < lower
    mid = ~(hi | lo)
    out[mid] = torch.acos(x[mid])
    out[hi] = math.acos(upper) + (x[hi] - upper) * _dacos(upper)
    out[lo] = math.acos(lower) + (x[lo] - lower) * _dacos(lower)
    return out


def so3_rotation_angle(R, eps: float = 1e-4, cos_angle: bool = False, cos_bound: float = 1e-4):
    """pytorch3d.transforms.so3.so3_rotation_angle: angle from the trace, ValueError outside [-1-eps, 3+eps]."""
    rot_trace = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    if ((rot_trace < -1.0 - eps) + (rot_trace > 3.0 + eps)).any():
        raise ValueError("A matrix has trace outside valid range [-1-eps,3+eps].")
    phi_cos = (rot_trace - 1.0) * 0.5
    if cos_angle:
        return phi_cos
    if cos_bound > 0.0:
        return acos_linear_extrapolation(phi_cos, (-1.0 + cos_bound, 1.0 - cos_bound))
    return torch.acos(phi_cos)


def so3_relative_angle(R1, R2, cos_angle: bool = False, cos_bound: float = 1e-4, eps: float = 1e-4):
    """pytorch3d.transforms.so3.so3_relative_angle: rotation angle of R1 R2^T."""
    return so3_rotation_angle(torch.bmm(R1, R2.permute(0, 2, 1)), cos_angle=cos_angle, cos_bound=cos_bound, eps=eps)
