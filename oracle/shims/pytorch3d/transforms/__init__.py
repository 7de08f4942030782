from .rotation_conversions import matrix_to_quaternion, quaternion_to_matrix
from .so3 import acos_linear_extrapolation, hat, so3_relative_angle, so3_rotation_angle


def _absent(*args, **kwargs):
    raise NotImplementedError("not on the sampling hot path; name exists for import only")


se3_exp_map = se3_log_map = _absent


class Transform3d:  # import-time name only (pose_diffusion_model.py:25)
    pass


class Rotate:
    pass


class Translate:
    pass
