"""Tensor-level binding of libgi2d: the replacement for the reference's pybind11 module.

Every function here has the NAME, ARGUMENT ORDER and RETURN SHAPE of the `m.def` it replaces in
the reference (`gsplat/gsplat/cuda/csrc/ext.cpp:4-69`, bodies in `csrc/bindings.cu`), so the
Python operators above it (gaussianimage_plus_b200/gsplat/*.py) read like the reference's.  The
difference is underneath: outputs are allocated here as torch tensors (caching allocator, graph
pools) and handed to the C ABI as raw pointers together with torch's CURRENT stream -- the
reference launches on the legacy default stream (`bindings.cu:594`).

Error behaviour mirrors `csrc/bindings.h:9-14` (CHECK_INPUT): non-CUDA or non-contiguous inputs
raise RuntimeError with the same wording; C-ABI failures raise Gi2dError (a RuntimeError).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib

TILE = 16


def _check_input(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    try:   # (the happy path first: these checks run ~30 times per iteration of a host-bound loop)
        if t.is_cuda and t.is_contiguous() and (dtype is None or t.dtype == dtype):
            return t
    except AttributeError:
        raise TypeError(f"{name} must be a torch.Tensor") from None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    raise RuntimeError(f"{name} must have dtype {dtype}, got {t.dtype}")


def _p(t: Optional[torch.Tensor]):
    # (a plain int converts to a c_void_p argument as well, without building a ctypes object)
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(dev) -> C.c_void_p:
    if _raw_stream is not None:
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        return C.c_void_p(_raw_stream(idx))
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


_WS_CACHE = {}


def _workspace(nbytes: int, dev) -> torch.Tensor:
    """Scratch for one C-ABI call.  Cached per (device, stream): calls on one stream execute in order, so the
    next call may reuse the bytes; a fresh torch.empty per call costs ~3 us of host time, and the operator path
    is host bound (several calls per iteration)."""
    nbytes = max(int(nbytes), 256)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(dev).cuda_stream)
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = _WS_CACHE[key] = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
    return ws


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NULL = _NullCtx()


def _on(dev):
    """`with _on(dev):` == `with _on(dev):` without the ~5 us of the context manager when `dev`
    already is the current device (the common case)."""
    idx = dev.index
    if idx is None or idx == torch.cuda.current_device():
        return _NULL
    return torch.cuda.device(dev)


f32, i32, i64 = torch.float32, torch.int32, torch.int64


# ------------------------------------------------------------------------------- projection
def _project_fwd(fn_name, num_points, clip_coe, means2d, p3, rot, img_height, img_width, tile_bounds,
                 radius_clip):
    lib = _lib.load()
    _check_input(means2d, "means2d", f32)
    _check_input(p3, "L_elements", f32)
    dev = means2d.device
    n = int(num_points)
    xys = torch.empty((n, 2), dtype=f32, device=dev)
    depths = torch.empty((n,), dtype=f32, device=dev)
    radii = torch.empty((n,), dtype=i32, device=dev)
    conics = torch.empty((n, 3), dtype=f32, device=dev)
    nth = torch.empty((n,), dtype=i32, device=dev)
    args = [n, _p(means2d), _p(p3)]
    if rot is not None:
        _check_input(rot, "rotation", f32)
        args.append(_p(rot))
    args += [int(img_width), int(img_height), int(tile_bounds[0]), int(tile_bounds[1]), float(clip_coe),
             float(radius_clip), _p(xys), _p(depths), _p(radii), _p(conics), _p(nth), _stream(dev)]
    with _on(dev):
        _lib.check(getattr(lib, fn_name)(*args), fn_name)
    return xys, depths, radii, conics, nth


def project_gaussians_2d_covariance_forward(num_points, clip_coe, means2d, L_elements, img_height, img_width,
                                            tile_bounds, clip_thresh, radius_clip, isprint=False):
    """ext.cpp:38 / bindings.cu:1455-1513"""
    return _project_fwd("gi2d_project_cov_fwd", num_points, clip_coe, means2d, L_elements, None, img_height,
                        img_width, tile_bounds, radius_clip)


def project_gaussians_2d_forward(num_points, clip_coe, means2d, L_elements, img_height, img_width, tile_bounds,
                                 clip_thresh, radius_clip, isprint=False):
    """ext.cpp:31 / bindings.cu:1317-1381"""
    return _project_fwd("gi2d_project_chol_fwd", num_points, clip_coe, means2d, L_elements, None, img_height,
                        img_width, tile_bounds, radius_clip)


def project_gaussians_2d_scale_rot_forward(num_points, clip_coe, means2d, scales2d, rotation, img_height,
                                           img_width, tile_bounds, clip_thresh, radius_clip, isprint=False):
    """ext.cpp:33 / bindings.cu:1384-1448"""
    return _project_fwd("gi2d_project_rs_fwd", num_points, clip_coe, means2d, scales2d, rotation, img_height,
                        img_width, tile_bounds, radius_clip)


def compute_cov2d_bounds(num_pts, clip_coe, cov2d):
    """ext.cpp:58 / bindings.cu:41-67 -> (conics [N,3], radii [N,1] float)"""
    lib = _lib.load()
    _check_input(cov2d, "covs2d", f32)
    dev = cov2d.device
    n = int(num_pts)
    conics = torch.empty((n, 3), dtype=f32, device=dev)
    radii = torch.empty((n, 1), dtype=f32, device=dev)
    with _on(dev):
        _lib.check(lib.gi2d_compute_cov2d_bounds(n, float(clip_coe), _p(cov2d), _p(conics), _p(radii), _stream(dev)),
                   "compute_cov2d_bounds")
    return conics, radii


def _bwd_common(radii, conics, v_xy, v_conic):
    _check_input(radii, "radii", i32)
    _check_input(conics, "conics", f32)
    v_xy = _check_input(v_xy.contiguous(), "v_xy", f32)
    v_conic = _check_input(v_conic.contiguous(), "v_conic", f32)
    return v_xy, v_conic


def project_gaussians_2d_covariance_backward(num_points, means2d, L_elements, img_height, img_width, radii, conics,
                                             v_xy, v_depth, v_conic):
    """ext.cpp:39 / bindings.cu:1565-1612 -> (v_cov2d, v_mean2d, v_L_elements)"""
    lib = _lib.load()
    v_xy, v_conic = _bwd_common(radii, conics, v_xy, v_conic)
    dev, n = conics.device, int(num_points)
    v_cov2d = torch.empty((n, 3), dtype=f32, device=dev)
    v_mean2d = torch.empty((n, 2), dtype=f32, device=dev)
    v_L = torch.empty((n, 3), dtype=f32, device=dev)
    with _on(dev):
        _lib.check(lib.gi2d_project_cov_bwd(n, _p(radii), _p(conics), _p(v_xy), _p(v_conic), _p(v_cov2d),
                                            _p(v_mean2d), _p(v_L), _stream(dev)), "project_cov_bwd")
    return v_cov2d, v_mean2d, v_L


def project_gaussians_2d_backward(num_points, means2d, L_elements, img_height, img_width, radii, conics, v_xy,
                                  v_depth, v_conic):
    """ext.cpp:32 / bindings.cu:1516-1562 -> (v_cov2d, v_mean2d, v_L_elements)"""
    lib = _lib.load()
    _check_input(L_elements, "L_elements", f32)
    v_xy, v_conic = _bwd_common(radii, conics, v_xy, v_conic)
    dev, n = conics.device, int(num_points)
    v_cov2d = torch.empty((n, 3), dtype=f32, device=dev)
    v_mean2d = torch.empty((n, 2), dtype=f32, device=dev)
    v_L = torch.empty((n, 3), dtype=f32, device=dev)
    with _on(dev):
        _lib.check(lib.gi2d_project_chol_bwd(n, _p(L_elements), int(img_width), int(img_height), _p(radii),
                                             _p(conics), _p(v_xy), _p(v_conic), _p(v_cov2d), _p(v_mean2d),
                                             _p(v_L), _stream(dev)), "project_chol_bwd")
    return v_cov2d, v_mean2d, v_L


def project_gaussians_2d_scale_rot_backward(num_points, means2d, scales2d, rotation, img_height, img_width, radii,
                                            conics, v_xy, v_depth, v_conic):
    """ext.cpp:34 / bindings.cu:1615-1668 -> (v_cov2d, v_mean2d, v_scale, v_rot)"""
    lib = _lib.load()
    _check_input(scales2d, "scales2d", f32)
    _check_input(rotation, "rotation", f32)
    v_xy, v_conic = _bwd_common(radii, conics, v_xy, v_conic)
    dev, n = conics.device, int(num_points)
    v_cov2d = torch.empty((n, 3), dtype=f32, device=dev)
    v_mean2d = torch.empty((n, 2), dtype=f32, device=dev)
    v_scale = torch.empty((n, 2), dtype=f32, device=dev)
    v_rot = torch.empty((n, 1), dtype=f32, device=dev)
    with _on(dev):
        _lib.check(lib.gi2d_project_rs_bwd(n, _p(scales2d), _p(rotation), _p(radii), _p(conics), _p(v_xy),
                                           _p(v_conic), _p(v_cov2d), _p(v_mean2d), _p(v_scale), _p(v_rot),
                                           _stream(dev)), "project_rs_bwd")
    return v_cov2d, v_mean2d, v_scale, v_rot


# ------------------------------------------------------------------------------- binning
def cumsum_i32(num_tiles_hit: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Replaces torch.cumsum(..., dtype=int32) (gsplat/utils.py:248).  Returns (cum, total[1]) on device."""
    lib = _lib.load()
    _check_input(num_tiles_hit, "num_tiles_hit", i32)
    dev, n = num_tiles_hit.device, num_tiles_hit.numel()
    cum = torch.empty((n,), dtype=i32, device=dev)
    total = torch.empty((1,), dtype=i32, device=dev)
    ws = _workspace(lib.gi2d_scan_workspace_size(n), dev)
    with _on(dev):
        _lib.check(lib.gi2d_cumsum_i32(n, _p(num_tiles_hit), _p(cum), _p(total), _p(ws), ws.numel(), _stream(dev)),
                   "cumsum_i32")
    return cum, total


def map_gaussian_to_intersects(num_points, num_intersects, xys, depths, radii, cum_tiles_hit, tile_bounds,
                               radius_clip=1.0, isprint=False):
    """ext.cpp:60 / bindings.cu:283-365 -> (isect_ids i64[I], gaussian_ids i32[I]), zero-initialised"""
    lib = _lib.load()
    _check_input(xys, "xys", f32)
    _check_input(depths, "depths", f32)
    _check_input(radii, "radii", i32)
    _check_input(cum_tiles_hit, "cum_tiles_hit", i32)
    dev = xys.device
    isect_ids = torch.zeros((int(num_intersects),), dtype=i64, device=dev)
    gaussian_ids = torch.zeros((int(num_intersects),), dtype=i32, device=dev)
    with _on(dev):
        _lib.check(lib.gi2d_map_gaussian_to_intersects(int(num_points), _p(xys), _p(depths), _p(radii),
                                                       _p(cum_tiles_hit), int(tile_bounds[0]), int(tile_bounds[1]),
                                                       float(radius_clip), _p(isect_ids), _p(gaussian_ids),
                                                       _stream(dev)), "map_gaussian_to_intersects")
    return isect_ids, gaussian_ids


def sort_pairs_i64(keys: torch.Tensor, vals: torch.Tensor, begin_bit: int = 0, end_bit: int = 64):
    """Replaces torch.sort + torch.gather (gsplat/utils.py:301-302): stable, signed, ascending."""
    lib = _lib.load()
    _check_input(keys, "isect_ids", i64)
    _check_input(vals, "gaussian_ids", i32)
    dev, n = keys.device, keys.numel()
    keys_out, vals_out = torch.empty_like(keys), torch.empty_like(vals)
    ws = _workspace(lib.gi2d_sort_workspace_size(n), dev)
    with _on(dev):
        _lib.check(lib.gi2d_sort_pairs_i64(n, _p(keys), _p(vals), _p(keys_out), _p(vals_out), int(begin_bit),
                                           int(end_bit), _p(ws), ws.numel(), _stream(dev)), "sort_pairs_i64")
    return keys_out, vals_out


def get_tile_bin_edges(num_intersects, isect_ids_sorted, num_rows: Optional[int] = None):
    """ext.cpp:66 / bindings.cu:368-383 -> tile_bins i32[rows,2]; rows = num_intersects like the reference
    unless `num_rows` is given (pass max(num_intersects, #tiles) to cover every tile)."""
    lib = _lib.load()
    _check_input(isect_ids_sorted, "isect_ids_sorted", i64)
    dev = isect_ids_sorted.device
    rows = int(num_intersects) if num_rows is None else int(num_rows)
    tile_bins = torch.empty((rows, 2), dtype=i32, device=dev)
    with _on(dev):
        _lib.check(lib.gi2d_get_tile_bin_edges(int(num_intersects), _p(isect_ids_sorted), _p(tile_bins), rows,
                                               _stream(dev)), "get_tile_bin_edges")
    return tile_bins


# ------------------------------------------------------------------------------- rasterize
_BIN_CAP = {}      # device index -> intersection capacity that has been enough so far
_BIN_PENDING = {}  # device index -> (pinned info, event) of the last gi2d_bin_sort call
_BIN_RING = {}     # device index -> [8 pinned i32[3] buffers, next]


class BinSortResult:
    """Outputs of `bin_sort`: the reference's (isect_ids_sorted, gaussian_ids_sorted, tile_bins) with `capacity`
    rows, `info` i32[3] on the device = {rows written, num_intersects, overflow}, and a pinned host copy of it that
    `check()` waits for (no synchronisation unless somebody asks)."""
    __slots__ = ("isect_ids_sorted", "gaussian_ids_sorted", "tile_bins", "info", "_host", "_event", "_dev", "capacity")

    def check(self, block: bool = True):
        """This call's counters: raises if the buffers overflowed (the render was built from a truncated list), else
        returns num_intersects.  The capacity for later calls grows with what has been seen.  `block=False`: only
        if the counters have already arrived (returns None otherwise) -- the autograd backward uses that, so the
        operator path never stalls the host; a late overflow is then caught by the next call."""
        if not block and not self._event.query():
            return None
        self._event.synchronize()
        rows, total, ovf = (int(v) for v in self._host)
        idx = self._dev
        _BIN_CAP[idx] = max(_BIN_CAP.get(idx, 0), 2 * total)
        if ovf:
            raise _lib.Gi2dError(f"bin_sort: {total} intersections did not fit the {self.capacity} rows allocated "
                                 "(the capacity has been raised; run the step again)")
        return total


def bin_sort(num_points, xys, depths, radii, tile_bounds, radius_clip=1.0) -> BinSortResult:
    """gi2d_bin_sort: compute_cumulative_intersects + bin_and_sort_gaussians (gsplat/utils.py:231-311) in ONE call with
    num_intersects kept on the device -- no `.item()`.  Requires depths that all share one bit pattern (the 2-D
    projections of this package emit 0.0 and say so)."""
    lib = _lib.load()
    _check_input(xys, "xys", f32)
    _check_input(radii, "radii", i32)
    dev = xys.device
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    n = int(num_points)
    tx, ty = int(tile_bounds[0]), int(tile_bounds[1])
    prev = _BIN_PENDING.get(idx)
    if prev is not None and prev[1].query():     # the previous call's counters have arrived: size from them
        _BIN_CAP[idx] = max(_BIN_CAP.get(idx, 0), 2 * int(prev[0][1]))
        if int(prev[0][2]):
            _BIN_PENDING.pop(idx, None)
            raise _lib.Gi2dError(f"bin_sort: the previous call overflowed its {prev[2]} rows ({int(prev[0][1])} "
                                 "intersections); the capacity has been raised -- run the step again")
    cap = int(min(max(_BIN_CAP.get(idx, 0), 1 << 16, 32 * n), max(n, 1) * tx * ty, 2 ** 31 - 1024))
    res = BinSortResult()
    res.isect_ids_sorted = torch.empty(cap, dtype=i64, device=dev)
    res.gaussian_ids_sorted = torch.empty(cap, dtype=i32, device=dev)
    res.tile_bins = torch.empty((tx * ty, 2), dtype=i32, device=dev)
    res.info = torch.empty(3, dtype=i32, device=dev)
    res.capacity, res._dev = cap, idx
    ws = _workspace(lib.gi2d_bin_sort_workspace_size(n, tx, ty, cap), dev)
    with _on(dev):
        _lib.check(lib.gi2d_bin_sort(n, _p(xys), _p(depths), _p(radii), tx, ty, float(radius_clip), cap,
                                     _p(res.isect_ids_sorted), _p(res.gaussian_ids_sorted), _p(res.tile_bins),
                                     _p(res.info), _p(ws), ws.numel(), _stream(dev)), "bin_sort")
        ring = _BIN_RING.get(idx)
        if ring is None:   # (NOT setdefault(idx, [...]): its argument would be built -- 8 pinned allocations -- on every call)
            ring = _BIN_RING[idx] = [[torch.zeros(3, dtype=i32).pin_memory() for _ in range(8)], 0]
        res._host = ring[0][ring[1] & 7]     # (pinned allocations cost ~50 us each: a small ring, reused)
        ring[1] += 1
        res._host.copy_(res.info, non_blocking=True)
        res._event = torch.cuda.Event()
        res._event.record(torch.cuda.current_stream(dev))
    _BIN_PENDING[idx] = (res._host, res._event, cap)
    return res


def rasterize_sum_plus_forward_dev(tile_bounds, block, img_size, bins: BinSortResult, xys, conics, colors, opacities,
                                   background):
    """rasterize_sum_plus_forward on the outputs of `bin_sort` (num_intersects on the device: the constant
    background image of rasterize_sum_plus.py:110-118 is the kernel's branch, not the host's)."""
    lib = _lib.load()
    _check_tiles(block)
    for t, nme, dt in ((xys, "xys", f32), (conics, "conics", f32), (colors, "colors", f32), (opacities, "opacities", f32)):
        _check_input(t, nme, dt)
    if colors.shape[-1] != 3:
        raise ValueError("rasterize_sum kernels render 3 channels")
    dev = xys.device
    W, H = int(img_size[0]), int(img_size[1])
    out_img = torch.empty((H, W, 3), dtype=f32, device=dev)
    final_Ts = torch.empty((H, W), dtype=f32, device=dev)
    bg = None
    if background is not None:
        bg = _check_input(background.to(device=dev, dtype=f32).contiguous(), "background", f32)
    with _on(dev):
        _lib.check(lib.gi2d_rasterize_sum_fwd_dev(int(tile_bounds[0]), int(tile_bounds[1]), W, H,
                                                  _p(bins.gaussian_ids_sorted), _p(bins.tile_bins),
                                                  int(bins.tile_bins.shape[0]), _p(xys), _p(conics), _p(colors),
                                                  _p(opacities), _p(out_img), _p(final_Ts), None,
                                                  _p(bins.info[1:2]), _p(bg), _stream(dev)), "rasterize_sum_fwd_dev")
    return out_img, final_Ts


def _check_tiles(block):
    if int(block[0]) != TILE or int(block[1]) != TILE:
        # the reference silently mis-bins for any other block size: refuse instead
        raise ValueError(f"only {TILE}x{TILE} tiles are supported (got block={tuple(block)})")


def rasterize_sum_plus_forward(tile_bounds, block, img_size, gaussian_ids_sorted, tile_bins, xys, conics, colors,
                               opacities, background=None, isprint=False):
    """ext.cpp:23 (and :16 `rasterize_sum_forward`) / bindings.cu:529-610
    -> (out_img [H,W,3], final_Ts [H,W], final_idx [H,W])"""
    lib = _lib.load()
    _check_tiles(block)
    for t, nme, dt in ((gaussian_ids_sorted, "gaussian_ids_sorted", i32), (tile_bins, "tile_bins", i32),
                       (xys, "xys", f32), (conics, "conics", f32), (colors, "colors", f32),
                       (opacities, "opacities", f32)):
        _check_input(t, nme, dt)
    if colors.shape[-1] != 3:
        raise ValueError("rasterize_sum kernels render 3 channels")
    dev = xys.device
    W, H = int(img_size[0]), int(img_size[1])
    out_img = torch.empty((H, W, 3), dtype=f32, device=dev)
    final_Ts = torch.empty((H, W), dtype=f32, device=dev)
    final_idx = torch.empty((H, W), dtype=i32, device=dev)
    with _on(dev):
        _lib.check(lib.gi2d_rasterize_sum_fwd(int(tile_bounds[0]), int(tile_bounds[1]), W, H,
                                              _p(gaussian_ids_sorted), _p(tile_bins), int(tile_bins.shape[0]),
                                              _p(xys), _p(conics), _p(colors), _p(opacities), _p(out_img),
                                              _p(final_Ts), _p(final_idx), _stream(dev)), "rasterize_sum_fwd")
    return out_img, final_Ts, final_idx


rasterize_sum_forward = rasterize_sum_plus_forward


def rasterize_sum_plus_backward(img_height, img_width, BLOCK_H, BLOCK_W, gaussian_ids_sorted, tile_bins, xys,
                                conics, colors, opacities, background, final_Ts, final_idx, v_output,
                                v_output_alpha=None):
    """ext.cpp:24 (and :17) / bindings.cu:1241-1314 -> (v_xy [N,2], v_conic [N,3], v_colors [N,3], v_opacity [N,1])"""
    lib = _lib.load()
    _check_tiles((BLOCK_W, BLOCK_H))
    for t, nme, dt in ((gaussian_ids_sorted, "gaussian_ids_sorted", i32), (tile_bins, "tile_bins", i32),
                       (xys, "xys", f32), (conics, "conics", f32), (colors, "colors", f32),
                       (opacities, "opacities", f32)):
        _check_input(t, nme, dt)
    v_output = _check_input(v_output.contiguous(), "v_output", f32)
    if xys.ndimension() != 2 or xys.size(1) != 2:
        raise RuntimeError("xys must have dimensions (num_points, 2)")  # bindings.cu:1269-1271
    if colors.ndimension() != 2 or colors.size(1) != 3:
        raise RuntimeError("colors must have 2 dimensions")  # bindings.cu:1273-1275
    dev, n = xys.device, xys.shape[0]
    H, W = int(img_height), int(img_width)
    tb = ((W + TILE - 1) // TILE, (H + TILE - 1) // TILE)
    # the four outputs back to back in ONE allocation: the library zero-fills them with one memset
    block = torch.empty(9 * n, dtype=f32, device=dev)
    v_xy, v_conic = block[:2 * n].view(n, 2), block[2 * n:5 * n].view(n, 3)
    v_colors, v_opacity = block[5 * n:8 * n].view(n, 3), block[8 * n:].view(n, 1)
    with _on(dev):
        _lib.check(lib.gi2d_rasterize_sum_bwd(n, tb[0], tb[1], W, H, _p(gaussian_ids_sorted), _p(tile_bins),
                                              int(tile_bins.shape[0]), _p(xys), _p(conics), _p(colors),
                                              _p(opacities), _p(final_idx), _p(v_output), _p(v_xy), _p(v_conic),
                                              _p(v_colors), _p(v_opacity), _stream(dev)), "rasterize_sum_bwd")
    return v_xy, v_conic, v_colors, v_opacity


rasterize_sum_backward = rasterize_sum_plus_backward


# ------------------------------------------------------------------------------- image losses
def image_loss_grad(render_hwc: torch.Tensor, gt_hwc: torch.Tensor, loss_type: str = "SSIM",
                    lambda_value: float = 0.7) -> Tuple[torch.Tensor, torch.Tensor]:
    """d loss_fn(clamp(render), gt, loss_type) / d render for the losses of models/utils.py:60-80 that
    have an SSIM term or are pointwise (L2, L1, SSIM, Fusion1-3), in two kernels (gi2d_loss.cu).
    render f32[H,W,3] (unclamped rasterizer output), gt f32 or u8 [H,W,3].
    Returns (v_out f32[H,W,3], ssim_sum f64[1]); mean SSIM = ssim_sum / (3 (H-10)(W-10))."""
    from .fit import loss_weights, msssim_term

    lib = _lib.load()
    _check_input(render_hwc, "render", f32)
    _check_input(gt_hwc, "gt")
    H, W, _ = render_hwc.shape
    w2, w1, ws = loss_weights(loss_type, lambda_value)
    wm, win = msssim_term(loss_type, lambda_value)
    dev = render_hwc.device
    v_out = torch.empty_like(render_hwc)
    ssim_sum = torch.zeros(1, dtype=torch.float64, device=dev)
    if wm:
        # Fusion4 / Fusion_hinerv (models/utils.py:76-79): the second return value is ms_ssim itself
        ws_buf = _workspace(lib.gi2d_msssim_grad_workspace_size(H, W), dev)
        is_u8 = gt_hwc.dtype == torch.uint8
        with _on(dev):
            _lib.check(lib.gi2d_image_msssim_loss_grad(H, W, win, _p(render_hwc), None if is_u8 else _p(gt_hwc),
                                                       _p(gt_hwc) if is_u8 else None, wm, w1 / (3.0 * H * W),
                                                       _p(v_out), _p(ssim_sum), _p(ws_buf), ws_buf.numel(),
                                                       _stream(dev)), "image_msssim_loss_grad")
        return v_out, ssim_sum
    ws_buf = _workspace(lib.gi2d_ssim_workspace_size(H, W), dev)
    is_u8 = gt_hwc.dtype == torch.uint8
    with _on(dev):
        _lib.check(lib.gi2d_image_loss_grad(H, W, _p(render_hwc), None if is_u8 else _p(gt_hwc),
                                            _p(gt_hwc) if is_u8 else None, ws, 2.0 * w2 / (3.0 * H * W),
                                            w1 / (3.0 * H * W), _p(v_out), _p(ssim_sum), _p(ws_buf), ws_buf.numel(),
                                            _stream(dev)), "image_loss_grad")
    return v_out, ssim_sum


MS_SSIM_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)   # pytorch_msssim.ms_ssim defaults


def ms_ssim(render_hwc: torch.Tensor, gt_hwc: torch.Tensor) -> float:
    """MS-SSIM of clamp(render) against the target (train.py:190: `ms_ssim(render, gt, data_range=1,
    size_average=True)`), evaluated by gi2d_ms_ssim.  render f32[H,W,3], gt f32 or u8 [H,W,3].  Synchronises."""
    lib = _lib.load()
    _check_input(render_hwc, "render", f32)
    _check_input(gt_hwc, "gt")
    H, W, _ = render_hwc.shape
    dev = render_hwc.device
    sums = torch.zeros(5, 3, 2, dtype=torch.float64, device=dev)
    ws = _workspace(lib.gi2d_ms_ssim_workspace_size(H, W), dev)
    is_u8 = gt_hwc.dtype == torch.uint8
    with _on(dev):
        _lib.check(lib.gi2d_ms_ssim(H, W, _p(render_hwc), None if is_u8 else _p(gt_hwc), _p(gt_hwc) if is_u8 else None,
                                    _p(sums), _p(ws), ws.numel(), _stream(dev)), "ms_ssim")
    s = sums.cpu()
    val = torch.ones(3, dtype=torch.float64)
    h, w = H, W
    for lvl in range(5):
        n = (h - 10) * (w - 10)
        mean = s[lvl, :, 1 if lvl < 4 else 0] / n          # cs for levels 0..3, ssim for the last
        val = val * torch.relu(mean) ** MS_SSIM_WEIGHTS[lvl]
        h, w = (h + 2 * (h & 1) - 2) // 2 + 1, (w + 2 * (w & 1) - 2) // 2 + 1
    return float(val.mean())
