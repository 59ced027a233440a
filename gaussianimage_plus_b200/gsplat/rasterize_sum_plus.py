"""Drop-in for gsplat/gsplat/rasterize_sum_plus.py:14-75 (the live model's rasterizer)."""
from typing import Optional

import torch
from torch import Tensor

from ._functions import RasterizeSum as _RasterizeGaussiansSum


def _prepare(xys, colors, background):
    if colors.dtype == torch.uint8:
        colors = colors.float() / 255  # rasterize_sum_plus.py:39-41
    if background is not None:
        assert background.shape[0] == colors.shape[-1], (
            f"incorrect shape of background color tensor, expected shape {colors.shape[-1]}")
    else:
        background = torch.ones(colors.shape[-1], dtype=torch.float32, device=colors.device)
    if xys.ndimension() != 2 or xys.size(1) != 2:
        raise ValueError("xys must have dimensions (N, 2)")
    if colors.ndimension() != 2:
        raise ValueError("colors must have dimensions (N, D)")
    return colors, background


def rasterize_gaussians_plus(xys: Tensor, depths: Tensor, radii: Tensor, conics: Tensor, num_tiles_hit: Tensor,
                             colors: Tensor, opacity: Tensor, img_height: int, img_width: int, BLOCK_H: int = 16,
                             BLOCK_W: int = 16, background: Optional[Tensor] = None,
                             return_alpha: Optional[bool] = False, radius_clip: float = 1.0,
                             isprint: bool = False) -> Tensor:
    """Accumulated-sum rasterization: out[H,W,C] = sum_g colors_g * min(1, opacity_g * exp(-sigma_g)), over the
    first 256 Gaussians of each 16x16 tile.  Differentiable w.r.t. xys, conics, colors, opacity."""
    colors, background = _prepare(xys, colors, background)
    return _RasterizeGaussiansSum.apply(xys.contiguous(), depths.contiguous(), radii.contiguous(),
                                        conics.contiguous(), num_tiles_hit.contiguous(), colors.contiguous(),
                                        opacity.contiguous(), img_height, img_width, BLOCK_H, BLOCK_W,
                                        background.contiguous(), radius_clip, isprint, getattr(depths, "_gi2d_depths_zero", None))
