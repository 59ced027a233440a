"""Drop-in for gsplat/gsplat/rasterize_sum.py:14-95.

The reference's 3-channel branch is broken as shipped (it passes 11 arguments to a 10-argument
binding and unpacks 4 results from 3, rasterize_sum.py:157-169 vs bindings.cu:457-526; SURVEY #13).
This module implements what its callers expect:

  clean form  (upstream GaussianImage; gaussianimage_cholesky.py:441, gaussianimage_rs.py:601)
      rasterize_gaussians_sum(xys, depths, radii, conics, num_tiles_hit, colors, opacity, H, W,
                              BLOCK_H, BLOCK_W, background=None, return_alpha=False) -> out_img
  stale form  (gaussianimage_cholesky.py:218, gaussianimage_rs.py:236): a `screenspace_points [N,4]`
      tensor as 2nd positional -> (out_img, cnt_gs_counts, screenspace_points)
"""
from typing import Optional

import torch
from torch import Tensor

from ._functions import RasterizeSum as _RasterizeGaussiansSum
from .rasterize_sum_plus import _prepare


def rasterize_gaussians_sum(xys: Tensor, *args, **kwargs):
    stale = len(args) > 0 and isinstance(args[0], Tensor) and args[0].dim() == 2 and args[0].shape[-1] == 4 \
        and args[0].shape[0] == xys.shape[0] and len(args) + len(kwargs) >= 9
    screenspace_points = None
    if stale:
        screenspace_points, args = args[0], args[1:]
    out = _sum(xys, *args, **kwargs)
    if not stale:
        return out
    img = out[0] if isinstance(out, tuple) else out
    # per-pixel contributor count: declared by the reference (bindings.cu:506-508), never filled
    cnt_gs_counts = torch.zeros(img.shape[0], img.shape[1], dtype=torch.int32, device=img.device)
    return img, cnt_gs_counts, screenspace_points


def _sum(xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width, BLOCK_H=16,
         BLOCK_W=16, background: Optional[Tensor] = None, return_alpha: Optional[bool] = False,
         isprint: bool = False):
    colors, background = _prepare(xys, colors, background)
    if colors.shape[-1] != 3:
        raise NotImplementedError("N-channel accumulate-sum (nd_rasterize_sum_*) is outside the hot path (SURVEY 2.3)")
    out_img = _RasterizeGaussiansSum.apply(xys.contiguous(), depths.contiguous(), radii.contiguous(),
                                           conics.contiguous(), num_tiles_hit.contiguous(), colors.contiguous(),
                                           opacity.contiguous(), img_height, img_width, BLOCK_H, BLOCK_W,
                                           background.contiguous(), 1.0, isprint,
                                           getattr(depths, "_gi2d_depths_zero", None))
    if return_alpha:
        # out_alpha = 1 - final_Ts and the kernel leaves T == 1 (forward.cu:617,682)
        return out_img, torch.zeros(img_height, img_width, device=out_img.device)
    return out_img
