"""Drop-in for gsplat/gsplat/project_gaussians_2d_scale_rot.py:12-66 (scale + rotation parameterisation)."""
from typing import Tuple

from torch import Tensor

from ._functions import ProjectScaleRot as _ProjectGaussians2dScaleRot


def project_gaussians_2d_scale_rot(means2d: Tensor, scales2d: Tensor, rotation: Tensor, img_height: int,
                                   img_width: int, tile_bounds: Tuple[int, int, int], clip_thresh: float = 0.01,
                                   coords_norm: bool = False, radius_clip: float = 1.0, isprint: bool = False):
    """(means2d [N,2] pixels, scales2d [N,2], rotation [N,1]) -> (xys, depths, radii, conics, num_tiles_hit)."""
    out = _ProjectGaussians2dScaleRot.apply(means2d.contiguous(), scales2d.contiguous(), rotation.contiguous(),
                                             img_height, img_width, tile_bounds, clip_thresh, radius_clip, isprint)
    out[1]._gi2d_depths_zero = True   # the 2-D projections emit depth 0.0: rasterize_* need not check
    return out
