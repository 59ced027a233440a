"""Autograd operators of the drop-in `gsplat` package (SURVEY 8b "Autograd contract").

One module holds the four torch.autograd.Function classes; the reference-named modules next to
it (project_gaussians_2d*.py, rasterize_sum*.py) expose them under the reference's signatures.
Saved tensors and returned gradient sets are the reference's:
  projection : saves (params, radii, conics)            -> grads for (means2d, params)
  rasterize  : saves (ids, bins, xys, conics, colors, opacity) -> grads for (xys, conics, colors, opacity)
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import cuda as _C
from .. import binding as _b
from .utils import bin_and_sort_gaussians, cumulative_intersects_and_depth_flag


class ProjectCovariance(Function):
    """reference: gsplat/gsplat/project_gaussians_2d_covariance.py:66-159"""

    @staticmethod
    def forward(ctx, means2d, cov2d, img_height, img_width, tile_bounds, clip_thresh, clip_coe, radius_clip,
                isprint):
        n = means2d.shape[-2]
        out = _C.project_gaussians_2d_covariance_forward(n, clip_coe, means2d, cov2d, img_height, img_width,
                                                         tile_bounds, clip_thresh, radius_clip, isprint)
        xys, depths, radii, conics, num_tiles_hit = out
        ctx.dims = (n, img_height, img_width)
        ctx.save_for_backward(means2d, cov2d, radii, conics)
        ctx.mark_non_differentiable(depths, radii, num_tiles_hit)
        return out

    @staticmethod
    def backward(ctx, v_xys, v_depths, v_radii, v_conics, v_num_tiles_hit):
        means2d, cov2d, radii, conics = ctx.saved_tensors
        n, H, W = ctx.dims
        _, v_mean2d, v_cov = _C.project_gaussians_2d_covariance_backward(
            n, means2d, cov2d, H, W, radii, conics, _dense(v_xys, (n, 2), means2d),
            None, _dense(v_conics, (n, 3), means2d))
        return (v_mean2d, v_cov) + (None,) * 7


class ProjectCholesky(Function):
    """reference: gsplat/gsplat/project_gaussians_2d.py:63-153 (clip_coe fixed at 3.0, :88)"""

    @staticmethod
    def forward(ctx, means2d, L_elements, img_height, img_width, tile_bounds, clip_thresh, radius_clip, isprint):
        n = means2d.shape[-2]
        out = _C.project_gaussians_2d_forward(n, 3.0, means2d, L_elements, img_height, img_width, tile_bounds,
                                              clip_thresh, radius_clip, isprint)
        xys, depths, radii, conics, num_tiles_hit = out
        ctx.dims = (n, img_height, img_width)
        ctx.save_for_backward(means2d, L_elements, radii, conics)
        ctx.mark_non_differentiable(depths, radii, num_tiles_hit)
        return out

    @staticmethod
    def backward(ctx, v_xys, v_depths, v_radii, v_conics, v_num_tiles_hit):
        means2d, L_elements, radii, conics = ctx.saved_tensors
        n, H, W = ctx.dims
        _, v_mean2d, v_L = _C.project_gaussians_2d_backward(
            n, means2d, L_elements, H, W, radii, conics, _dense(v_xys, (n, 2), means2d), None,
            _dense(v_conics, (n, 3), means2d))
        return (v_mean2d, v_L) + (None,) * 6


class ProjectScaleRot(Function):
    """reference: gsplat/gsplat/project_gaussians_2d_scale_rot.py:69-166 (clip_coe fixed at 3.0, :94)"""

    @staticmethod
    def forward(ctx, means2d, scales2d, rotation, img_height, img_width, tile_bounds, clip_thresh, radius_clip,
                isprint):
        n = means2d.shape[-2]
        if n < 1 or means2d.shape[-1] != 2:
            raise ValueError(f"Invalid shape for means2d: {means2d.shape}")
        out = _C.project_gaussians_2d_scale_rot_forward(n, 3.0, means2d, scales2d, rotation, img_height, img_width,
                                                        tile_bounds, clip_thresh, radius_clip, isprint)
        xys, depths, radii, conics, num_tiles_hit = out
        ctx.dims = (n, img_height, img_width)
        ctx.save_for_backward(means2d, scales2d, rotation, radii, conics)
        ctx.mark_non_differentiable(depths, radii, num_tiles_hit)
        return out

    @staticmethod
    def backward(ctx, v_xys, v_depths, v_radii, v_conics, v_num_tiles_hit):
        means2d, scales2d, rotation, radii, conics = ctx.saved_tensors
        n, H, W = ctx.dims
        _, v_mean2d, v_scale, v_rot = _C.project_gaussians_2d_scale_rot_backward(
            n, means2d, scales2d, rotation, H, W, radii, conics, _dense(v_xys, (n, 2), means2d), None,
            _dense(v_conics, (n, 3), means2d))
        return (v_mean2d, v_scale, v_rot.view_as(rotation)) + (None,) * 6


def _dense(g, shape, like):
    """autograd hands None for outputs that did not take part in the graph"""
    if g is None:
        return torch.zeros(shape, dtype=torch.float32, device=like.device)
    return g.contiguous()


class RasterizeSum(Function):
    """reference: gsplat/gsplat/rasterize_sum_plus.py:78-243 (== rasterize_sum.py:97-343 for 3 channels)"""

    @staticmethod
    def forward(ctx, xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width, BLOCK_H,
                BLOCK_W, background, radius_clip, isprint, depths_zero=None):
        n = xys.size(0)
        tile_bounds = ((img_width + BLOCK_W - 1) // BLOCK_W, (img_height + BLOCK_H - 1) // BLOCK_H, 1)
        if depths_zero:
            # depths straight from one of this package's 2-D projections (known to be all 0.0): ONE binning call
            # with num_intersects kept on the device (gi2d_bin_sort) -- no .item(), no torch.sort -- and the
            # rasterizer takes the "no intersection" branch of rasterize_sum_plus.py:110-118 by itself
            res = _b.bin_sort(n, xys, depths.view(-1), radii.view(-1), tile_bounds, radius_clip)
            out_img, final_Ts = _b.rasterize_sum_plus_forward_dev(
                tile_bounds, (BLOCK_W, BLOCK_H, 1), (img_width, img_height, 1), res, xys, conics, colors, opacity,
                background)
            ctx.meta = (img_height, img_width, BLOCK_H, BLOCK_W, 1)
            ctx.bins = res
            ctx.save_for_backward(res.gaussian_ids_sorted, res.tile_bins, xys, conics, colors, opacity)
            ctx.final_Ts = final_Ts
            return out_img
        else:
            num_intersects, cum_tiles_hit, uniform = cumulative_intersects_and_depth_flag(num_tiles_hit, depths)
        if num_intersects < 1:
            # rasterize_sum_plus.py:110-118 -- a constant background image, no gradients
            out_img = torch.ones(img_height, img_width, colors.shape[-1], device=xys.device) * background
            ids = torch.zeros(0, dtype=torch.int32, device=xys.device)
            bins = torch.zeros(0, 2, dtype=torch.int32, device=xys.device)
            final_Ts = torch.zeros(img_height, img_width, device=xys.device)
        else:
            _, _, _, ids, bins = bin_and_sort_gaussians(n, num_intersects, xys, depths, radii, cum_tiles_hit,
                                                        tile_bounds, radius_clip, _depths_uniform=uniform)
            out_img, final_Ts, _ = _C.rasterize_sum_plus_forward(
                tile_bounds, (BLOCK_W, BLOCK_H, 1), (img_width, img_height, 1), ids, bins, xys, conics, colors,
                opacity, background, isprint)
        ctx.meta = (img_height, img_width, BLOCK_H, BLOCK_W, num_intersects)
        ctx.bins = None
        ctx.save_for_backward(ids, bins, xys, conics, colors, opacity)
        ctx.final_Ts = final_Ts
        return out_img

    @staticmethod
    def backward(ctx, v_out_img):
        H, W, BH, BW, num_intersects = ctx.meta
        ids, bins, xys, conics, colors, opacity = ctx.saved_tensors
        if ctx.bins is not None:
            ctx.bins.check(block=False)   # (raises if the forward worked on a truncated list; never waits)
        if num_intersects < 1:
            grads = tuple(torch.zeros_like(t) for t in (xys, conics, colors, opacity))
        else:
            grads = _C.rasterize_sum_plus_backward(H, W, BH, BW, ids, bins, xys, conics, colors, opacity, None,
                                                   None, None, v_out_img.contiguous(), None)
        v_xy, v_conic, v_colors, v_opacity = grads
        return (v_xy, None, None, v_conic, None, v_colors, v_opacity.view_as(opacity)) + (None,) * 8
