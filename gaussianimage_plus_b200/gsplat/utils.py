"""Binning helpers of the drop-in `gsplat` package (reference: gsplat/gsplat/utils.py:12-311).
Same names, argument order and return tuples; cumsum / sort / gather are libgi2d kernels instead
of torch.cumsum / torch.sort / torch.gather."""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from . import cuda as _C
from .. import binding as _b


def cumulative_intersects_and_depth_flag(num_tiles_hit: Tensor, depths: Tensor):
    """(num_intersects, cum_tiles_hit, depths_uniform) with ONE host read-back: the total of the prefix sum and
    "all depths share one bit pattern" travel together (the reference's .item() at utils.py:249 is the one
    synchronisation this path keeps)."""
    cum, total = _b.cumsum_i32(num_tiles_hit.contiguous().view(-1).to(torch.int32))
    if depths.numel() == 0:
        return int(total.item()), cum, True
    lo, hi = torch.aminmax(depths.contiguous().view(-1).view(torch.int32))
    both = torch.stack((total[0], (lo == hi).to(torch.int32))).tolist()
    return int(both[0]), cum, bool(both[1])


def compute_cumulative_intersects(num_tiles_hit: Tensor) -> Tuple[int, Tensor]:
    """utils.py:231-250: (int num_intersects, cum_tiles_hit int32[N]); reads one int back like the reference."""
    cum, total = _b.cumsum_i32(num_tiles_hit.contiguous().view(-1).to(torch.int32))
    return int(total.item()), cum


def map_gaussian_to_intersects(num_points, num_intersects, xys, depths, radii, cum_tiles_hit, tile_bounds,
                               radius_clip=1.0, isprint=False) -> Tuple[Tensor, Tensor]:
    """utils.py:12-57"""
    return _C.map_gaussian_to_intersects(num_points, num_intersects, xys.contiguous(),
                                         depths.contiguous().view(-1), radii.contiguous().view(-1),
                                         cum_tiles_hit.contiguous(), tile_bounds, radius_clip, isprint)


def get_tile_bin_edges(num_intersects, isect_ids_sorted, num_rows=None) -> Tensor:
    """utils.py:166-187.  `num_rows` (extension) sizes tile_bins; default = num_intersects as the reference."""
    return _C.get_tile_bin_edges(num_intersects, isect_ids_sorted.contiguous(), num_rows)


def compute_cov2d_bounds(cov2d: Tensor, clip_coe: float = 3.0) -> Tuple[Tensor, Tensor]:
    """utils.py:190-209"""
    assert cov2d.shape[-1] == 3, (
        f"Expected input cov2d to be of shape (*batch, 3) (upper triangular values), but got {tuple(cov2d.shape)}")
    num_pts = cov2d.shape[0]
    assert num_pts > 0
    return _C.compute_cov2d_bounds(num_pts, clip_coe, cov2d.contiguous())


def bin_and_sort_gaussians(num_points, num_intersects, xys, depths, radii, cum_tiles_hit, tile_bounds,
                           radius_clip=1.0, isprint=False, _depths_uniform=None):
    """utils.py:253-311 -> (isect_ids, gaussian_ids, isect_ids_sorted, gaussian_ids_sorted, tile_bins).

    tile_bins gets max(num_intersects, #tiles) rows so that every tile has a row (the
    reference allocates num_intersects rows and reads past them when there are fewer)."""
    depths = depths.contiguous().view(-1)
    isect_ids, gaussian_ids = map_gaussian_to_intersects(num_points, num_intersects, xys, depths, radii,
                                                         cum_tiles_hit, tile_bounds, radius_clip, isprint)
    begin_bit, end_bit = 0, 64
    if depths.numel() > 0:
        # when every depth has the same bit pattern (the 2-D projections emit 0.0) only the tile bits order keys
        # (`_depths_uniform`: the caller already knows -- rasterize_gaussians_* read the answer back together
        #  with num_intersects, one synchronisation instead of two)
        if _depths_uniform is None:
            lo, hi = torch.aminmax(depths.view(torch.int32))
            _depths_uniform = bool(lo == hi)
        if _depths_uniform:
            tiles = max(int(tile_bounds[0]) * int(tile_bounds[1]), 2)
            begin_bit, end_bit = 32, min(64, 32 + (tiles - 1).bit_length())
    isect_ids_sorted, gaussian_ids_sorted = _b.sort_pairs_i64(isect_ids, gaussian_ids, begin_bit, end_bit)
    rows = max(int(num_intersects), int(tile_bounds[0]) * int(tile_bounds[1]))
    tile_bins = get_tile_bin_edges(num_intersects, isect_ids_sorted, rows)
    return isect_ids, gaussian_ids, isect_ids_sorted, gaussian_ids_sorted, tile_bins
