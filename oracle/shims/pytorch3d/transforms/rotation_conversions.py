import torch


def quaternion_to_matrix(quaternions):
    """Real-first (w,x,y,z); scale 2/|q|^2 so any non-zero q maps to a rotation."""
    w, x, y, z = torch.unbind(quaternions, -1)
    s2 = 2.0 / (quaternions * quaternions).sum(-1)
    rows = (
        1 - s2 * (y * y + z * z), s2 * (x * y - z * w), s2 * (x * z + y * w),
        s2 * (x * y + z * w), 1 - s2 * (x * x + z * z), s2 * (y * z - x * w),
        s2 * (x * z - y * w), s2 * (y * z + x * w), 1 - s2 * (x * x + y * y),
    )
    return torch.stack(rows, -1).reshape(quaternions.shape[:-1] + (3, 3))


def matrix_to_quaternion(matrix):
    raise NotImplementedError("training-only (camera_transform.py:115)")
