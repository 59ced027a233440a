"""Drop-in for gsplat/gsplat/project_gaussians_2d.py:12-60 (Cholesky parameterisation)."""
from typing import Tuple

from torch import Tensor

from ._functions import ProjectCholesky as _ProjectGaussians2d


def project_gaussians_2d(means2d: Tensor, L_elements: Tensor, img_height: int, img_width: int,
                         tile_bounds: Tuple[int, int, int], clip_thresh: float = 0.01, radius_clip: float = 1.0,
                         isprint: bool = False):
    """(means2d [N,2] in [-1,1], L_elements [N,3] = (l11, l21, l22)) -> (xys, depths, radii, conics, num_tiles_hit).
    The 3-sigma extent (clip_coe = 3.0) is fixed as in the reference (project_gaussians_2d.py:88)."""
    out = _ProjectGaussians2d.apply(means2d.contiguous(), L_elements.contiguous(), img_height, img_width,
                                     tile_bounds, clip_thresh, radius_clip, isprint)
    out[1]._gi2d_depths_zero = True   # the 2-D projections emit depth 0.0: rasterize_* need not check
    return out
