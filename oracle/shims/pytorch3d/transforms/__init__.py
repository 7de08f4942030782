from .rotation_conversions import matrix_to_quaternion, quaternion_to_matrix
from .so3 import hat


def _absent(*args, **kwargs):
    raise NotImplementedError("not on the sampling hot path; name exists for import only")


se3_exp_map = se3_log_map = so3_relative_angle = _absent


class Transform3d:  # import-time name only (pose_diffusion_model.py:25)
    pass


class Rotate:
    pass


class Translate:
    pass
