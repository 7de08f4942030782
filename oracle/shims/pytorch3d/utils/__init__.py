import torch


def opencv_from_cameras_projection(cameras, image_size):
    """PyTorch3D NDC camera -> OpenCV (R, tvec, K).  image_size rows are (H, W).

    Flip the x/y axes (first two columns of R, first two entries of T), transpose R,
    K = [[fx*s, 0, W/2 - px*s], [0, fy*s, H/2 - py*s], [0, 0, 1]], s = min(H, W)/2.
    """
    R_ndc = cameras.R.clone()
    T_ndc = cameras.T.clone()
    T_ndc[:, :2] *= -1
    R_ndc[:, :, :2] *= -1
    R_cv = R_ndc.permute(0, 2, 1)
    size_wh = image_size.to(R_cv).flip(dims=(1,))
    half_min = size_wh.min(dim=1, keepdim=True).values / 2.0
    half_min = half_min.expand(-1, 2)
    centre = size_wh / 2.0
    K = torch.zeros_like(R_cv)
    K[:, :2, 2] = centre - cameras.principal_point * half_min
    K[:, 2, 2] = 1.0
    fl = cameras.focal_length * half_min
    K[:, 0, 0] = fl[:, 0]
    K[:, 1, 1] = fl[:, 1]
    return R_cv, T_ndc, K
