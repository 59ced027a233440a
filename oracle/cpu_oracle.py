"""TEST INFRASTRUCTURE ONLY -- numpy front-end of the CPU oracle (oracle/gi2d_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module; the product package (gaussianimage_plus_b200) never does.

Every function takes/returns numpy arrays with the reference's layouts and dtypes and follows the
reference source cited in gi2d_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "gi2d_oracle.c")
_SO = os.path.join(_HERE, "_build", "libgi2d_oracle.so")
_lib = None

TILE = 16


def build(force: bool = False) -> str:
    """gcc -O2 -fopenmp -ffp-contract=off (explicit fmaf only) -> oracle/_build/libgi2d_oracle.so"""
    if force or not os.path.exists(_SO) or (
        os.path.exists(_SRC) and os.path.getmtime(_SRC) > os.path.getmtime(_SO)
    ):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
             "-o", _SO, _SRC, "-lm"]
        )
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_cumsum_i32.restype = C.c_int32
        _lib.orc_fit_step.restype = C.c_double
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def tile_bounds(H, W):
    return ((W + TILE - 1) // TILE, (H + TILE - 1) // TILE, 1)


def _project(fn, means2d, p3, rot, H, W, tb, clip_coe, radius_clip):
    n = means2d.shape[0]
    xys = np.zeros((n, 2), np.float32)
    depths = np.zeros((n,), np.float32)
    radii = np.zeros((n,), np.int32)
    conics = np.zeros((n, 3), np.float32)
    nth = np.zeros((n,), np.int32)
    args = [C.c_int(n), _p(means2d), _p(p3)]
    if rot is not None:
        args.append(_p(rot))
    args += [C.c_int(W), C.c_int(H), C.c_int(tb[0]), C.c_int(tb[1]), C.c_float(clip_coe),
             C.c_float(radius_clip), _p(xys), _p(depths), _p(radii), _p(conics), _p(nth)]
    fn(*args)
    return xys, depths, radii, conics, nth


def project_cov_fwd(means2d, cov2d, H, W, tb=None, clip_coe=3.0, radius_clip=1.0):
    means2d, cov2d = _f(means2d), _f(cov2d)
    return _project(lib().orc_project_cov_fwd, means2d, cov2d, None, H, W, tb or tile_bounds(H, W),
                    clip_coe, radius_clip)


def project_chol_fwd(means2d, L, H, W, tb=None, clip_coe=3.0, radius_clip=1.0):
    means2d, L = _f(means2d), _f(L)
    return _project(lib().orc_project_chol_fwd, means2d, L, None, H, W, tb or tile_bounds(H, W),
                    clip_coe, radius_clip)


def project_rs_fwd(means2d, scales, rot, H, W, tb=None, clip_coe=3.0, radius_clip=1.0):
    means2d, scales, rot = _f(means2d), _f(scales), _f(rot).reshape(-1)
    return _project(lib().orc_project_rs_fwd, means2d, scales, rot, H, W, tb or tile_bounds(H, W),
                    clip_coe, radius_clip)


def compute_cov2d_bounds(cov2d, clip_coe=3.0):
    cov2d = _f(cov2d)
    n = cov2d.shape[0]
    conics = np.zeros((n, 3), np.float32)
    radii = np.zeros((n,), np.float32)
    lib().orc_compute_cov2d_bounds(C.c_int(n), C.c_float(clip_coe), _p(cov2d), _p(conics), _p(radii))
    return conics, radii


def project_bwd(mode, p3, rot, H, W, radii, conics, v_xy, v_conic):
    """mode 0 covariance / 1 Cholesky / 2 scale-rot -> (v_cov2d, v_mean2d, v_p3, v_rot|None)"""
    n = conics.shape[0]
    radii, conics, v_xy, v_conic = _i(radii), _f(conics), _f(v_xy), _f(v_conic)
    p3 = None if p3 is None else _f(p3)
    rot = None if rot is None else _f(rot).reshape(-1)
    w3 = 2 if mode == 2 else 3
    v_cov2d = np.zeros((n, 3), np.float32)
    v_mean = np.zeros((n, 2), np.float32)
    v_p3 = np.zeros((n, w3), np.float32)
    v_rot = np.zeros((n,), np.float32) if mode == 2 else None
    lib().orc_project_bwd(C.c_int(mode), C.c_int(n), _p(p3), _p(rot), C.c_int(W), C.c_int(H), _p(radii),
                          _p(conics), _p(v_xy), _p(v_conic), _p(v_cov2d), _p(v_mean), _p(v_p3), _p(v_rot))
    return v_cov2d, v_mean, v_p3, v_rot


def cumsum_i32(num_tiles_hit):
    a = _i(num_tiles_hit)
    out = np.zeros_like(a)
    total = lib().orc_cumsum_i32(C.c_int(a.shape[0]), _p(a), _p(out)) if a.shape[0] else 0
    return int(total), out


def map_gaussian_to_intersects(num_intersects, xys, depths, radii, cum, tb, radius_clip=1.0):
    xys, depths, radii, cum = _f(xys), _f(depths).reshape(-1), _i(radii), _i(cum)
    ids = np.zeros((num_intersects,), np.int64)
    gids = np.zeros((num_intersects,), np.int32)
    lib().orc_map_gaussian_to_intersects(C.c_int(xys.shape[0]), _p(xys), _p(depths), _p(radii), _p(cum),
                                         C.c_int(tb[0]), C.c_int(tb[1]), C.c_float(radius_clip), _p(ids),
                                         _p(gids))
    return ids, gids


def sort_pairs(keys, vals):
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    vals = _i(vals)
    ko, vo = np.zeros_like(keys), np.zeros_like(vals)
    lib().orc_sort_pairs_i64(C.c_int(keys.shape[0]), _p(keys), _p(vals), _p(ko), _p(vo))
    return ko, vo


def get_tile_bin_edges(sorted_ids, rows):
    sorted_ids = np.ascontiguousarray(sorted_ids, dtype=np.int64)
    bins = np.zeros((rows, 2), np.int32)
    lib().orc_get_tile_bin_edges(C.c_int(sorted_ids.shape[0]), _p(sorted_ids), _p(bins), C.c_int(rows))
    return bins


def bin_and_sort(xys, depths, radii, nth, tb, radius_clip=1.0, rows=None):
    """utils.py:231-311 composed.  rows defaults to #tiles."""
    total, cum = cumsum_i32(nth)
    ids, gids = map_gaussian_to_intersects(total, xys, depths, radii, cum, tb, radius_clip)
    ids_s, gids_s = sort_pairs(ids, gids)
    bins = get_tile_bin_edges(ids_s, tb[0] * tb[1] if rows is None else rows)
    return total, cum, ids, gids, ids_s, gids_s, bins


def rasterize_sum_fwd(H, W, gids_sorted, tile_bins, xys, conics, colors, opacities=None, with_slack=False):
    tb = tile_bounds(H, W)
    gids_sorted, tile_bins = _i(gids_sorted), _i(tile_bins)
    xys, conics, colors = _f(xys), _f(conics), _f(colors)
    opac = None if opacities is None else _f(opacities).reshape(-1)
    out = np.zeros((H, W, 3), np.float32)
    Ts = np.zeros((H, W), np.float32)
    fidx = np.zeros((H, W), np.int32)
    slack = np.zeros((H, W), np.float32) if with_slack else None
    lib().orc_rasterize_sum_fwd(C.c_int(tb[0]), C.c_int(tb[1]), C.c_int(W), C.c_int(H), _p(gids_sorted),
                                _p(tile_bins), C.c_int(tile_bins.shape[0]), _p(xys), _p(conics), _p(colors),
                                _p(opac), _p(out), _p(Ts), _p(fidx), _p(slack))
    return (out, Ts, fidx, slack) if with_slack else (out, Ts, fidx)


def rasterize_sum_bwd(H, W, gids_sorted, tile_bins, xys, conics, colors, opacities, v_output, with_slack=False):
    tb = tile_bounds(H, W)
    n = xys.shape[0]
    gids_sorted, tile_bins = _i(gids_sorted), _i(tile_bins)
    xys, conics, colors, v_output = _f(xys), _f(conics), _f(colors), _f(v_output)
    opac = None if opacities is None else _f(opacities).reshape(-1)
    v_xy = np.zeros((n, 2), np.float32)
    v_conic = np.zeros((n, 3), np.float32)
    v_colors = np.zeros((n, 3), np.float32)
    v_opacity = np.zeros((n,), np.float32)
    slack = np.zeros((n, 9), np.float32) if with_slack else None
    mag = np.zeros((n, 9), np.float32) if with_slack else None
    lib().orc_rasterize_sum_bwd(C.c_int(n), C.c_int(tb[0]), C.c_int(tb[1]), C.c_int(W), C.c_int(H),
                                _p(gids_sorted), _p(tile_bins), C.c_int(tile_bins.shape[0]), _p(xys),
                                _p(conics), _p(colors), _p(opac), _p(v_output), _p(v_xy), _p(v_conic),
                                _p(v_colors), _p(v_opacity), _p(slack), _p(mag))
    res = (v_xy, v_conic, v_colors, v_opacity)
    return res + (slack, mag) if with_slack else res


class FitCfg(C.Structure):
    _fields_ = [("n", C.c_int), ("img_w", C.c_int), ("img_h", C.c_int), ("clip_coe", C.c_float),
                ("radius_clip", C.c_float), ("lr0", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("lr_step_size", C.c_int), ("lr_gamma", C.c_float), ("step", C.c_int),
                ("color_sigmoid", C.c_int)]


class FitState:
    """In-place CPU restatement of GaussianImage_Covariance.train_iter (L2, Adam, StepLR)."""

    def __init__(self, xyz, cov, cov_bound, rgb, gt_hwc, lr=0.018, clip_coe=3.0, radius_clip=1.0,
                 color_sigmoid=False):
        self.xyz, self.cov, self.cov_bound, self.rgb = (np.array(_f(a)) for a in (xyz, cov, cov_bound, rgb))
        self.gt = _f(gt_hwc)
        self.H, self.W = self.gt.shape[:2]
        self.m = [np.zeros_like(a) for a in (self.xyz, self.cov, self.rgb)]
        self.v = [np.zeros_like(a) for a in (self.xyz, self.cov, self.rgb)]
        self.step = 0
        self.cfg = FitCfg(self.xyz.shape[0], self.W, self.H, clip_coe, radius_clip, lr, 0.9, 0.999, 1e-15,
                          20000, 0.5, 0, int(color_sigmoid))
        self.num_intersects = 0

    def train_iter(self, want_image=False):
        self.step += 1
        self.cfg.step = self.step
        img = np.zeros((self.H, self.W, 3), np.float32) if want_image else None
        ni = C.c_int32(0)
        mse = lib().orc_fit_step(C.byref(self.cfg), _p(self.xyz), _p(self.cov), _p(self.cov_bound), _p(self.rgb),
                                 _p(self.m[0]), _p(self.v[0]), _p(self.m[1]), _p(self.v[1]), _p(self.m[2]),
                                 _p(self.v[2]), _p(self.gt), _p(img), C.byref(ni))
        self.num_intersects = ni.value
        return (mse, img) if want_image else mse


# ---------------------------------------------------------------------------------------------------------
# The two operations that change the model's size, restated in numpy.  float32 arithmetic in
# torch's operation order; test infrastructure like everything else in this module.
def check_non_semi_definite(cov2d):
    """models/gaussianimage_covariance.py:373-382: valid = (c0*c2 - c1**2 > 0) & (c0 > 0) & (c2 > 0), float32.
    Returns (number to prune, valid mask)."""
    c = np.asarray(cov2d, np.float32)
    det = (c[:, 0] * c[:, 2]).astype(np.float32) - (c[:, 1] * c[:, 1]).astype(np.float32)
    valid = (det > 0) & (c[:, 0] > 0) & (c[:, 2] > 0)
    return int((~valid).sum()), valid


def non_semi_definite_prune(xyz, cov, rgb, bound, m, v):
    """models/gaussianimage_covariance.py:336-371 (`_prune_optimizer` + `non_semi_definite_prune`): when some but
    not all Gaussians fail the test on get_cov2d_elements (= _cov2d + cholesky_bound, :169), parameters, both Adam
    moments (dicts xyz / cov2d / f_dc) and the bound rows are masked together, order kept.
    Returns (to_prune, xyz, cov, rgb, bound, m, v)."""
    n_bad, valid = check_non_semi_definite(np.asarray(cov, np.float32) + np.asarray(bound, np.float32))
    if n_bad and xyz.shape[0] - n_bad > 0:
        return (n_bad, xyz[valid], cov[valid], rgb[valid], bound[valid],
                {k: t[valid] for k, t in m.items()}, {k: t[valid] for k, t in v.items()})
    return n_bad, xyz, cov, rgb, bound, m, v


def add_sample_positions(errors, cur_num_points, max_num_points, new_cov2d_draw, W, H, base_num_samples=1000,
                         last=False, slv=True):
    """train.py:85-118 + models/gaussianimage_covariance.py:307-334.  `errors` f32[H,W] = |render - gt| summed over
    the channels (train.py:87); k = max - cur when `last` (iter == iterations - grow_iter) else min(1000, max - cur)
    (:93-97); the k pixels of largest error, descending (torch.topk, :101; ties -- which torch leaves unspecified
    -- by the smaller index); new_xyz = (idx % W, idx // W) as float (:104-108), colour 0 (:106), covariance =
    the caller's `torch.rand(k, 3) + (0.5, 0, 0.5)` draw (:110-112; passed in so that both sides use the same
    numbers); candidates whose covariance fails check_non_semi_definite are dropped (:309-314); the appended rows
    get the bound (lp, 0, lp) with lp = min(H*W / (9 pi n_new), 300) of the NEW count (:326-330).
    (Dividing the errors by their sum, :89, does not change the order.)
    Returns (k, chosen pixel indices [k], valid mask [k], new_xyz, new_cov, new_bound rows of the valid ones)."""
    import math

    room = max(0, max_num_points - cur_num_points)
    k = room if last else min(base_num_samples, room)
    e = np.asarray(errors, np.float32).reshape(-1)
    order = np.lexsort((np.arange(e.size), -e.astype(np.float64)))[:k]   # value descending, index ascending
    cov = np.asarray(new_cov2d_draw, np.float32)[:k]
    _, valid = check_non_semi_definite(cov)
    idx = order[valid]
    new_xyz = np.stack([idx % W, idx // W], axis=1).astype(np.float32)
    n_new = cur_num_points + int(valid.sum())
    lp = np.float32(min(H * W / (9 * math.pi * n_new), 300)) if slv else np.float32(0.5)
    new_bound = np.tile(np.array([lp, 0, lp], np.float32), (int(valid.sum()), 1))
    return k, order, valid, new_xyz, cov[valid], new_bound
