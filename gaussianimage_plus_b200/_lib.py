"""ctypes binding of libgi2d.so (the C ABI declared in include/gi2d.h).

There is deliberately NO fallback: if the CUDA library is missing or a call fails, the product
path raises.  (The CPU oracle under oracle/ is test infrastructure and is never imported here.)
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libgi2d.so")

GI2D_OK = 0
ERR_NAMES = {-1: "GI2D_ERR_INVALID", -2: "GI2D_ERR_CUDA", -3: "GI2D_ERR_WORKSPACE"}

_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_SZ = C.c_size_t


class FitParams(C.Structure):
    """struct gi2d_fit_params (include/gi2d.h)"""
    _fields_ = [
        ("num_points", C.c_int32), ("img_width", C.c_int32), ("img_height", C.c_int32),
        ("tiles_x", C.c_int32), ("tiles_y", C.c_int32),
        ("tile_row_begin", C.c_int32), ("tile_row_end", C.c_int32),
        ("isect_capacity", C.c_int32),
        ("clip_coe", C.c_float), ("radius_clip", C.c_float),
        ("lr0", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("lr_step_size", C.c_int32), ("lr_gamma", C.c_float),
        ("color_sigmoid", C.c_int32), ("loss_scale", C.c_float), ("external_optimizer", C.c_int32),
        ("loss_l1_scale", C.c_float), ("loss_ssim_weight", C.c_float), ("dynamic_points", C.c_int32),
        ("loss_msssim_weight", C.c_float), ("loss_msssim_win", C.c_int32),
    ]


class FitBuffers(C.Structure):
    """struct gi2d_fit_buffers (include/gi2d.h)"""
    _fields_ = [
        ("xyz", _P), ("cov", _P), ("cov_bound", _P), ("rgb", _P),
        ("m_xyz", _P), ("v_xyz", _P), ("m_cov", _P), ("v_cov", _P), ("m_rgb", _P), ("v_rgb", _P),
        ("gt_hwc", _P), ("out_img", _P),
        ("grads", _P), ("proj", _P), ("sorted_keys", _P), ("tile_bins", _P), ("stats", _P),
        ("workspace", _P), ("workspace_bytes", _SZ), ("gt_u8_hwc", _P),
        ("best", _P), ("err_map", _P), ("best_bound", _P),
    ]


MAX_RANKS = 8


class TileRow(C.Structure):
    """struct gi2d_tilerow (include/gi2d.h)"""
    _fields_ = [
        ("rank", C.c_int32), ("world", C.c_int32), ("band_edge", C.c_int32 * (MAX_RANKS + 1)),
        ("own_begin", C.c_int32), ("own_end", C.c_int32), ("sync", C.c_int32),
        ("peer_grads", _P * MAX_RANKS), ("peer_proj", _P * MAX_RANKS), ("peer_boxes", _P * MAX_RANKS),
        ("peer_flags", _P * MAX_RANKS), ("ctrl", _P),
    ]


class QuantParams(C.Structure):
    """struct gi2d_quant_params (include/gi2d.h)"""
    _fields_ = [
        ("num_points", C.c_int32), ("xy_qmax", C.c_int32), ("cov_qmax", C.c_int32), ("color_qmax", C.c_int32),
        ("color_sigmoid", C.c_int32), ("lr0", C.c_float), ("lr_step", C.c_int32), ("lr_q0", C.c_float),
        ("lr_q_step", C.c_int32), ("lr_gamma", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
        ("eps", C.c_float), ("eps_xyz_q", C.c_float),
    ]


class QuantBuffers(C.Structure):
    """struct gi2d_quant_buffers (include/gi2d.h)"""
    _fields_ = [
        ("xyz", _P), ("cov", _P), ("rgb", _P), ("bound", _P),
        ("m_xyz", _P), ("v_xyz", _P), ("m_cov", _P), ("v_cov", _P), ("m_rgb", _P), ("v_rgb", _P),
        ("qparams", _P), ("qm", _P), ("qv", _P), ("qstats", _P),
        ("out_xyz", _P), ("out_cov", _P), ("out_rgb", _P), ("in_grads", _P), ("dbg_grads", _P),
    ]


# name -> (restype, argtypes); every symbol include/gi2d.h declares
SIGNATURES = {
    "gi2d_abi_version": (_I, []),
    "gi2d_last_error": (C.c_char_p, []),
    "gi2d_project_cov_fwd": (_I, [_I, _P, _P, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P]),
    "gi2d_project_chol_fwd": (_I, [_I, _P, _P, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P]),
    "gi2d_project_rs_fwd": (_I, [_I, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P]),
    "gi2d_compute_cov2d_bounds": (_I, [_I, _F, _P, _P, _P, _P]),
    "gi2d_project_cov_bwd": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gi2d_project_chol_bwd": (_I, [_I, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gi2d_project_rs_bwd": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gi2d_scan_workspace_size": (_SZ, [_I]),
    "gi2d_cumsum_i32": (_I, [_I, _P, _P, _P, _P, _SZ, _P]),
    "gi2d_map_gaussian_to_intersects": (_I, [_I, _P, _P, _P, _P, _I, _I, _F, _P, _P, _P]),
    "gi2d_sort_workspace_size": (_SZ, [_I]),
    "gi2d_sort_pairs_i64": (_I, [_I, _P, _P, _P, _P, _I, _I, _P, _SZ, _P]),
    "gi2d_get_tile_bin_edges": (_I, [_I, _P, _P, _I, _P]),
    "gi2d_bin_sort_workspace_size": (_SZ, [_I, _I, _I, _I]),
    "gi2d_bin_sort": (_I, [_I, _P, _P, _P, _I, _I, _F, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "gi2d_rasterize_sum_fwd": (_I, [_I, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gi2d_rasterize_sum_fwd_dev": (_I, [_I, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gi2d_rasterize_sum_bwd": (_I, [_I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gi2d_fit_workspace_size": (_SZ, [C.POINTER(FitParams)]),
    "gi2d_fit_forward_backward": (_I, [C.POINTER(FitParams), C.POINTER(FitBuffers), _I, _P]),
    "gi2d_fit_adam": (_I, [C.POINTER(FitParams), C.POINTER(FitBuffers), _P]),
    "gi2d_fit_launch_count": (_I, [C.POINTER(FitParams), _I]),
    "gi2d_fit_profile": (_I, [C.POINTER(FitParams), C.POINTER(FitBuffers), C.POINTER(C.c_float), _P]),
    "gi2d_fit_profile_raster": (_I, [C.POINTER(FitParams), C.POINTER(FitBuffers), _I, C.POINTER(C.c_float), _P]),
    "gi2d_measure_fp32_peak": (_I, [C.POINTER(C.c_float), _P]),
    "gi2d_fit_bucket_capacity": (_I, [C.POINTER(FitParams)]),
    "gi2d_fit_export_binning": (_I, [C.POINTER(FitParams), C.POINTER(FitBuffers), _P, _P, _P]),
    "gi2d_fit_input_grads": (_I, [C.POINTER(FitParams), C.POINTER(FitBuffers), _P, _P]),
    "gi2d_quant_init": (_I, [C.POINTER(QuantParams), C.POINTER(QuantBuffers), _P]),
    "gi2d_quant_forward": (_I, [C.POINTER(QuantParams), C.POINTER(QuantBuffers), _P]),
    "gi2d_quant_backward_step": (_I, [C.POINTER(QuantParams), C.POINTER(QuantBuffers), _P]),
    "gi2d_tilerow_init": (_I, [C.POINTER(FitParams), C.POINTER(FitBuffers), C.POINTER(TileRow), _P]),
    "gi2d_tilerow_step": (_I, [C.POINTER(FitParams), C.POINTER(FitBuffers), C.POINTER(TileRow), _I, _I, _P]),
    "gi2d_ssim_workspace_size": (_SZ, [_I, _I]),
    "gi2d_image_loss_grad": (_I, [_I, _I, _P, _P, _P, _F, _F, _F, _P, _P, _P, _SZ, _P]),
    "gi2d_msssim_grad_workspace_size": (_SZ, [_I, _I]),
    "gi2d_image_msssim_loss_grad": (_I, [_I, _I, _I, _P, _P, _P, _F, _F, _P, _P, _P, _SZ, _P]),
    "gi2d_ms_ssim_workspace_size": (_SZ, [_I, _I]),
    "gi2d_ms_ssim": (_I, [_I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "gi2d_host_pipe_create": (_I, [C.POINTER(_P)]),
    "gi2d_host_pipe_destroy": (_I, [_P]),
    "gi2d_fit_step_host": (_I, [_P, C.POINTER(FitParams), C.POINTER(FitBuffers), _P, _SZ, _P, _P, C.POINTER(_I)]),
    "gi2d_host_pipe_wait": (_I, [_P, _I]),
    "gi2d_fit_reset": (_I, [C.POINTER(FitParams), C.POINTER(FitBuffers), _I, _P]),
    "gi2d_fit_set_num_points": (_I, [C.POINTER(FitParams), C.POINTER(FitBuffers), _I, _P]),
    "gi2d_fit_prune_workspace_size": (_SZ, [_I]),
    "gi2d_fit_prune": (_I, [C.POINTER(FitParams), C.POINTER(FitBuffers), _P, _SZ, _P]),
    "gi2d_fit_densify_workspace_size": (_SZ, [_I, _I]),
    "gi2d_fit_densify": (_I, [C.POINTER(FitParams), C.POINTER(FitBuffers), _I, _P, _I, _P, _SZ, _P]),
}

_lib = None


class Gi2dError(RuntimeError):
    pass


def load():
    """Load libgi2d.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Gi2dError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C gaussianimage_plus_b200/csrc`.  There is no CPU fallback."
            )
        # GI2D_LIB: a differently-compiled libgi2d (tools/build_variant.sh) for A/B timing; still no fallback
        lib = C.CDLL(os.environ.get("GI2D_LIB") or LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != GI2D_OK:
        msg = load().gi2d_last_error().decode(errors="replace")
        raise Gi2dError(f"{what or 'gi2d'} failed: {ERR_NAMES.get(rc, rc)}: {msg}")
