"""Drop-in for gsplat/gsplat/project_gaussians_2d_covariance.py:11-63 (the live model's projection)."""
from typing import Tuple

from torch import Tensor

from ._functions import ProjectCovariance as _ProjectGaussians2d_covariance


def project_gaussians_2d_covariance(means2d: Tensor, L_elements: Tensor, img_height: int, img_width: int,
                                    tile_bounds: Tuple[int, int, int], clip_thresh: float = 0.01,
                                    coords_norm: bool = False, isprint: bool = False, clip_coe: float = 3.0,
                                    radius_clip: float = 1.0):
    """(means2d [N,2] pixels, L_elements [N,3] = (sxx, sxy, syy)) -> (xys, depths, radii, conics, num_tiles_hit).

    Differentiable w.r.t. means2d and L_elements.  `coords_norm`, `clip_thresh` are accepted and unused
    and `isprint` only enabled debug printing in the reference."""
    out = _ProjectGaussians2d_covariance.apply(means2d.contiguous(), L_elements.contiguous(), img_height,
                                                img_width, tile_bounds, clip_thresh, clip_coe, radius_clip, isprint)
    out[1]._gi2d_depths_zero = True   # the 2-D projections emit depth 0.0: rasterize_* need not check
    return out
