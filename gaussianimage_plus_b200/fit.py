"""Fused fit loop: the host-side mirror of `GaussianImage_Covariance` for the hot path.

`GaussianImageFitter` keeps the reference model's vocabulary (models/gaussianimage_covariance.py):
`_xyz`, `_cov2d`, `_features_dc`, `cholesky_bound`, `forward()`, `train_iter()`,
`densification_postfix()`, `non_semi_definite_prune()` -- but one `train_iter` is a single CUDA
graph replay of 2 kernels (gi2d_fit.cu) instead of ~70 launches and two host syncs
(SURVEY 3.1).  Nothing here computes on the CPU; without libgi2d.so it raises.

Two things differ from a literal transcription, both invisible in the results:

* PSNR is not read back every iteration: `train_iter()` only enqueues work; `psnr()` / `stats()`
  synchronise when the caller wants the number (the reference's per-iteration `.item()` at
  gaussianimage_covariance.py:257 is exactly the sync this removes).
* The optimiser step of iteration k is applied by the first kernel of iteration k+1 (the thread that
  projects Gaussian g first applies Adam to it).  Until then the gradient is *pending*; any host
  access to the parameters (`_xyz`, `_cov2d`, `_features_dc`, `exp_avg*`), `forward()`, pruning and
  densification flush it first, so the host always observes the reference's post-`optimizer.step()`
  values.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Tuple

import torch

from . import _lib
from .binding import TILE, _stream

STAT_STEP, STAT_ISECTS, STAT_OVERFLOW, STAT_LR, STAT_SSE, STAT_SSE_SLOTS = 0, 1, 2, 3, 16, 64
STAT_BEST_SSE, STAT_BEST_STEP, STAT_NON_PSD, STAT_SSIM_SUM, STAT_ABS_SUM = 9, 10, 11, 13, 14
STAT_NUM_POINTS, STAT_BEST_N, STAT_PRUNED, STAT_ADDED = 80, 81, 82, 83
STAT_MAX_TILE = 84
STAT_MSSSIM = 89
STAT_COUNT = 96
_NAMES = ("xyz", "cov2d", "f_dc")  # the reference's optimiser group names (gaussianimage_covariance.py:93-96)


def slv_bound(H: int, W: int, num_points: int) -> float:
    """SLV low-pass bound of the reference init: min(HW / (9 pi N), 300) (gaussianimage_covariance.py:61)."""
    return min(H * W / (9 * math.pi * num_points), 300)


# loss_fn of models/utils.py:60-80 as (w_mse, w_l1, w_ssim); lambda_value = 0.7 at both call sites
# (gaussianimage_covariance.py:222,252).  The MS-SSIM variants come back from msssim_term().
def loss_weights(loss_type: str, lambda_value: float = 0.7) -> Tuple[float, float, float]:
    lam = float(lambda_value)
    table = {"L2": (1.0, 0.0, 0.0), "L1": (0.0, 1.0, 0.0), "SSIM": (0.0, 0.0, 1.0),
             "Fusion1": (lam, 0.0, 1.0 - lam), "Fusion2": (0.0, lam, 1.0 - lam), "Fusion3": (lam, 1.0 - lam, 0.0),
             "Fusion4": (0.0, lam, 0.0), "Fusion_hinerv": (0.0, lam, 0.0)}
    if loss_type not in table:
        raise ValueError(f"loss_type {loss_type!r} is not supported (have {sorted(table)})")
    return table[loss_type]


def msssim_term(loss_type: str, lambda_value: float = 0.7) -> Tuple[float, int]:
    """(weight, win_size) of the `1 - ms_ssim` term: Fusion4 = 0.7 l1 + 0.3 (1 - ms_ssim), Fusion_hinerv the same
    with win_size = 5 (models/utils.py:76-79); (0, 11) for every other loss."""
    if loss_type == "Fusion4":
        return 1.0 - float(lambda_value), 11
    if loss_type == "Fusion_hinerv":
        return 1.0 - float(lambda_value), 5
    return 0.0, 11


def _sse_total(s) -> float:
    """Sum of the 64 squared-error partials in the order the kernels use for their best-so-far decision
    (sse_total_warp in gi2d_fit.cu: slot i + slot 32+i, then an xor butterfly), so that the `sse` of the best
    step and `best_sse` are the same double."""
    t = s[STAT_SSE:STAT_SSE + STAT_SSE_SLOTS].tolist()
    v = [t[i] + t[32 + i] for i in range(32)]
    for d in (16, 8, 4, 2, 1):
        v = [v[i] + v[i ^ d] for i in range(d)]   # only lanes < d feed lane 0
    return v[0]


def _flushed(attr):
    """Property over a private capacity-sized tensor (or dict of tensors): the getter first applies a pending
    optimiser step and returns the LIVE rows [0, cur_num_points) as a view, so in-place edits reach the device
    arrays; the setter copies into the live rows (or swaps the model in when the row count differs)."""
    def get(self):
        self.sync_params()
        n = self.cur_num_points
        v = getattr(self, attr)
        return {k: t[:n] for k, t in v.items()} if isinstance(v, dict) else v[:n]

    def set_(self, value):
        cur = getattr(self, attr, None)
        if cur is None or isinstance(cur, dict) or isinstance(value, dict):
            setattr(self, attr, value)
            return
        n = self.cur_num_points
        if value.shape[0] != n:
            raise ValueError(f"{attr}: {value.shape[0]} rows for a model of {n} Gaussians (use _replace())")
        self.sync_params()
        cur[:n].copy_(value)

    return property(get, set_)


class GaussianImageFitter:
    _xyz = _flushed("_t_xyz")                # f32[N,2] pixel coordinates
    _cov2d = _flushed("_t_cov2d")            # f32[N,3] covariance parameters (before the SLV bound)
    _features_dc = _flushed("_t_f_dc")       # f32[N,3] colours
    exp_avg = _flushed("_t_m")               # Adam first moments, dict keyed like the reference's groups
    exp_avg_sq = _flushed("_t_v")            # Adam second moments
    cholesky_bound = _flushed("_t_bound")    # f32[N,3] SLV bound added to the covariance parameters

    def __init__(self, num_points: int, H: int, W: int, device="cuda:0", lr: float = 0.018,
                 clip_coe: float = 3.0, radius_clip: float = 1.0, color_norm: bool = False,
                 SLV_init: bool = True, tile_rows: Optional[Tuple[int, int]] = None,
                 isect_capacity: Optional[int] = None, use_graph: bool = True,
                 grad_hook=None, loss_type: str = "L2", lambda_value: float = 0.7,
                 max_num_points: Optional[int] = None):
        """`max_num_points`: rows to allocate for (the model can grow to it by densification with no
        reallocation and no new CUDA graph); default: num_points."""
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.Gi2dError("GaussianImageFitter needs a CUDA device: there is no CPU path")
        self.H, self.W = int(H), int(W)
        self.tile_bounds = ((self.W + TILE - 1) // TILE, (self.H + TILE - 1) // TILE, 1)
        self.tile_rows = tile_rows if tile_rows is not None else (0, self.tile_bounds[1])
        self.lr, self.clip_coe, self.radius_clip = lr, clip_coe, radius_clip
        self.color_norm, self.SLV = bool(color_norm), bool(SLV_init)
        self.use_graph = use_graph
        self.loss_type, self.loss_w = loss_type, loss_weights(loss_type, lambda_value)
        self.loss_ms = msssim_term(loss_type, lambda_value)
        self.grad_hook = grad_hook  # called right after the backward of every step (multi-GPU all-reduce)
        self._capacity_hint = isect_capacity
        self._dirty = False         # a gradient is pending on the device
        self.external_optimizer = 0  # 1: the caller applies the gradients (codec.py); 2: parallel.TileRowFit
        self.keep_render = False    # tests: also store the unclamped [H,W,3] render of every train_iter
        self.track_best = True      # keep the parameters of the best-PSNR step on the device (train.py:132-137)
        f = dict(dtype=torch.float32, device=self.device)
        # The per-Gaussian arrays hold `capacity` rows; the LIVE count is kept ON THE DEVICE (stats slot
        # NUM_POINTS: prune / densify change it there) and mirrored on the host lazily (`cur_num_points`).
        self.capacity = max(int(num_points), int(max_num_points or 0))
        self._n, self._n_stale = int(num_points), False
        cap, n = self.capacity, self._n
        # reference init, gaussianimage_covariance.py:52-66
        w_init = torch.rand(n, 1, **f) * self.W
        h_init = torch.rand(n, 1, **f) * self.H
        self._t_xyz = torch.zeros(cap, 2, **f)
        self._t_xyz[:n] = torch.cat((w_init, h_init), dim=1)
        self._t_cov2d = torch.zeros(cap, 3, **f)
        self._t_cov2d[:n] = torch.rand((n, 3), **f)
        self._t_f_dc = torch.zeros(cap, 3, **f)
        lp = slv_bound(self.H, self.W, n) if self.SLV else 0.5
        self._t_bound = torch.zeros(cap, 3, **f)
        self._t_bound[:n] = torch.tensor([lp, 0, lp], **f)
        self.gt_hwc = None
        self._step0 = 0
        self._expected_step = 0
        self._alloc_state(zero_moments=True)

    # ------------------------------------------------------------------ buffers
    @property
    def cur_num_points(self) -> int:
        """The live Gaussian count.  After a device-side prune / densify the host's copy is stale: the first read
        fetches it (one 8-byte read-back; synchronises)."""
        if self._n_stale:
            self._n = int(self.stats_buf[STAT_NUM_POINTS].item())
            self._n_stale = False
        return self._n

    def _raw_params(self):
        return {"xyz": self._t_xyz, "cov2d": self._t_cov2d, "f_dc": self._t_f_dc}

    def _alloc_state(self, zero_moments: bool, eager_steps: int = 1):
        """(Re)allocate everything that is sized by the capacity or the intersection capacity."""
        n = self.capacity
        f = dict(dtype=torch.float32, device=self.device)
        if zero_moments:
            self._t_m = {k: torch.zeros_like(t) for k, t in self._raw_params().items()}
            self._t_v = {k: torch.zeros_like(t) for k, t in self._raw_params().items()}
        tiles = self.tile_bounds[0] * self.tile_bounds[1]
        # default: 32 intersections per Gaussian, and -- for the bucketed binning, whose tiles each own capacity / #tiles
        # rows -- at least 512 per tile, twice what the rasterizer stages (densification clusters new Gaussians: tiles
        # with 150-300 entries were seen at 768x512; a full bucket costs a regrow and the re-run of the lost iterations)
        cap = self._capacity_hint or max(1 << 16, 32 * n, int(os.environ.get("GI2D_ROWS_PER_TILE", "512")) * tiles)
        self.isect_capacity = int(min(cap, max(n, 1) * tiles, 2 ** 31 - 1024))
        if not getattr(self, "_keep_exchange_buffers", False):   # (parallel.TileRowFit homes them in peer memory)
            self.grads = torch.zeros(n, 8, **f)
            self.proj = torch.zeros(n, 8, **f)
        if getattr(self, "best", None) is None or self.best.shape[0] != n:
            old, oldb = getattr(self, "best", None), getattr(self, "best_bound", None)
            self.best = torch.zeros(n, 8, **f)          # device-side best-state snapshot (train.py:132-137)
            self.best_bound = torch.zeros(n, 3, **f)    # ... and its slv_bound (train.py:136)
            if old is not None:                          # (capacity grew: the snapshot stays valid)
                k = min(old.shape[0], n)
                self.best[:k], self.best_bound[:k] = old[:k], oldb[:k]
        if not hasattr(self, "err_map"):
            self.err_map = torch.zeros(self.H, self.W, **f)
        self._keys_buf = torch.zeros(self.isect_capacity, dtype=torch.int64, device=self.device)
        self._bins_buf = torch.zeros(tiles, 2, dtype=torch.int32, device=self.device)
        if not hasattr(self, "stats_buf"):
            self.stats_buf = torch.zeros(STAT_COUNT, dtype=torch.float64, device=self.device)
            self.stats_buf[STAT_NUM_POINTS] = float(self._n)
        self.out_hwc = torch.zeros(self.H, self.W, 3, **f)
        self.render_chw = torch.zeros(3, self.H, self.W, **f)
        self.params = _lib.FitParams(
            n, self.W, self.H, self.tile_bounds[0], self.tile_bounds[1], self.tile_rows[0], self.tile_rows[1],
            self.isect_capacity, self.clip_coe, self.radius_clip, self.lr, 0.9, 0.999, 1e-15, 20000, 0.5,
            int(self.color_norm), 2.0 * self.loss_w[0] / (3.0 * self.H * self.W),
            int(self.external_optimizer), self.loss_w[1] / (3.0 * self.H * self.W), self.loss_w[2], 1,
            self.loss_ms[0], self.loss_ms[1])
        ws_bytes = self.lib.gi2d_fit_workspace_size(C.byref(self.params))
        # rows per tile of the bucketed binning (0: the scan + placement path is in use, include/gi2d.h)
        self.bucket_cap = int(self.lib.gi2d_fit_bucket_capacity(C.byref(self.params)))
        self.workspace = torch.zeros(max(ws_bytes, 256), dtype=torch.uint8, device=self.device)
        self._prune_ws = None
        self._densify_ws = None
        self._invalidate_graphs()
        # steps to run un-graphed before the next capture: 1 loads the kernels (lazy module loading cannot be
        # captured)
        self._eager_left = eager_steps
        self._dirty = False
        self._bind()

    def _bind(self, out_img=None, err_map=False):
        m, v = self._t_m, self._t_v
        if out_img is None and self.keep_render:
            out_img = self.out_hwc.data_ptr()
        gt = self.gt_hwc
        self.buffers = _lib.FitBuffers(
            self._t_xyz.data_ptr(), self._t_cov2d.data_ptr(), self._t_bound.data_ptr(), self._t_f_dc.data_ptr(),
            m["xyz"].data_ptr(), v["xyz"].data_ptr(), m["cov2d"].data_ptr(), v["cov2d"].data_ptr(),
            m["f_dc"].data_ptr(), v["f_dc"].data_ptr(),
            gt.data_ptr() if (gt is not None and gt.dtype == torch.float32) else None,
            out_img, self.grads.data_ptr(), self.proj.data_ptr(), self._keys_buf.data_ptr(),
            self._bins_buf.data_ptr(), self.stats_buf.data_ptr(), self.workspace.data_ptr(), self.workspace.numel(),
            gt.data_ptr() if (gt is not None and gt.dtype == torch.uint8) else None,
            self.best.data_ptr() if self.track_best else None,
            self.err_map.data_ptr() if err_map else None,
            self.best_bound.data_ptr() if self.track_best else None)

    def _export_binning(self):
        """The reference's view of the last forward's binning (utils.py:301-302 / forward.cu:211-233): one
        ascending key array i64[capacity] (tile << 32 | gaussian) and tile_bins i32[tiles,2].  With the bucketed
        layout they are compacted on demand (gi2d_fit_export_binning; synchronises)."""
        if not self.bucket_cap:
            return self._keys_buf, self._bins_buf
        keys, bins = torch.zeros_like(self._keys_buf), torch.zeros_like(self._bins_buf)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gi2d_fit_export_binning(C.byref(self.params), C.byref(self.buffers), keys.data_ptr(),
                                                        bins.data_ptr(), _stream(self.device)), "fit_export_binning")
        return keys, bins

    @property
    def sorted_keys(self) -> torch.Tensor:
        return self._export_binning()[0]

    @property
    def tile_bins(self) -> torch.Tensor:
        return self._export_binning()[1]

    def reset_stats(self, step: int = 0):
        """Zero the device-side statistics (drops a pending gradient) and set the Adam step counter."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gi2d_fit_reset(C.byref(self.params), C.byref(self.buffers), int(step),
                                               _stream(self.device)), "fit_reset")
        self._dirty = False
        self._expected_step = int(step)   # training steps requested so far (the device counts the ones that happened)

    def sync_params(self, force: bool = False):
        """Apply a pending optimiser step now (asynchronous; no-op when nothing is pending).  `force`
        launches the flush kernel regardless, for its non-positive-definite count (stats()['non_psd'])."""
        if self._dirty or force:
            self._dirty = False
            with torch.cuda.device(self.device):
                _lib.check(self.lib.gi2d_fit_adam(C.byref(self.params), C.byref(self.buffers), _stream(self.device)),
                           "fit_adam")

    # ------------------------------------------------------------------ target
    def set_target(self, gt_image: torch.Tensor, overlap: bool = False):
        """gt_image: float32 [1,3,H,W] (the reference's layout, utils.py:21-27) or [H,W,3]; or uint8 [H,W,3]
        -- the image as stored: the kernels then use u8/255, the value ToTensor would have produced, and
        the host->device copy is 4x smaller.  Copied (asynchronously from pinned memory) to HWC on device.
        `overlap`: upload into the OTHER of two device buffers on a copy stream, so that the transfer of the
        next step's target runs under the step in flight (a driver that feeds a new target every step)."""
        if gt_image.dim() == 4:
            gt_image = gt_image[0].permute(1, 2, 0)
        assert gt_image.shape == (self.H, self.W, 3), gt_image.shape
        assert gt_image.dtype in (torch.float32, torch.uint8), gt_image.dtype
        if self.gt_hwc is None or self.gt_hwc.dtype != gt_image.dtype:
            first = self.gt_hwc is None
            self.gt_hwc = torch.empty(self.H, self.W, 3, dtype=gt_image.dtype, device=self.device)
            self._invalidate_graphs()
            self._gt_alt = None
            self.gt_hwc.copy_(gt_image, non_blocking=True)
            self._bind()
            if first:
                self.reset_stats(self._step0)
            return
        if not overlap:
            self.gt_hwc.copy_(gt_image, non_blocking=True)
            return
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_gt_alt", None) is None:
            self._gt_alt = torch.empty_like(self.gt_hwc)
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._ev_read = {}      # buffer -> event recorded after the last step that read it
            self._ev_copied = {t.data_ptr(): torch.cuda.Event() for t in (self.gt_hwc, self._gt_alt)}
        nxt, cur = self._gt_alt, self.gt_hwc
        free = self._ev_read.get(nxt.data_ptr())
        with torch.cuda.stream(self._copy_stream):
            if free is not None:
                self._copy_stream.wait_event(free)
            nxt.copy_(gt_image, non_blocking=True)
            self._ev_copied[nxt.data_ptr()].record(self._copy_stream)
        main.wait_event(self._ev_copied[nxt.data_ptr()])
        # swap the buffers, their graphs and their bound argument blocks
        self._graphs[cur.data_ptr()] = (self._graph, self.buffers)
        self.gt_hwc, self._gt_alt = nxt, cur
        self._graph, bound = self._graphs.get(nxt.data_ptr(), (None, None))
        if bound is not None:
            self.buffers = bound
        else:
            self._bind()

    def _invalidate_graphs(self):
        self._graph = None
        self._graphs = {}
        self._multi_graphs = {}
        self._pipe_bound = None
        self._render_bound = None

    # ------------------------------------------------------------------ steps fed from host memory
    def step_from_host(self, host_img: torch.Tensor, host_stats: torch.Tensor) -> int:
        """One train_iter whose target comes from PINNED host memory and whose stats block goes back to pinned
        host memory, through ONE C call (gi2d_fit_step_host): upload on the library's copy stream into the
        other of two device buffers (under the step in flight), the step's kernels, the stats block's read-back.
        Asynchronous; returns the slot to hand to `wait_host_result`.  host_img: u8 or f32 [H,W,3];
        host_stats: f64[STAT_COUNT].  The fitter's device must be the current CUDA device."""
        pb = self._pipe_bound
        if pb is None or self._pipe_dtype != host_img.dtype:
            pb = self._setup_host_pipe(host_img, host_stats)
        self._pipe_idx ^= 1
        rc = self._step_host(self._pipe, self._params_ref, pb[self._pipe_idx], host_img.data_ptr(), self._pipe_bytes,
                             host_stats.data_ptr(), torch.cuda.current_stream().cuda_stream, self._slot_ref)
        if rc != 0:
            _lib.check(rc, "fit_step_host")
        self._dirty = True
        self._expected_step += 1
        return self._slot.value

    def _setup_host_pipe(self, host_img, host_stats):
        assert host_img.shape == (self.H, self.W, 3) and host_img.is_pinned() and host_stats.is_pinned()
        assert host_stats.dtype == torch.float64 and host_stats.numel() >= STAT_COUNT
        if self.grad_hook is not None or self.external_optimizer:
            raise RuntimeError("step_from_host does not run the multi-GPU exchange hook: use set_target + train_iter")
        if torch.cuda.current_device() != self.device.index:
            raise RuntimeError("step_from_host: make the fitter's device current (torch.cuda.set_device)")
        if getattr(self, "_pipe", None) is None:
            self._pipe = C.c_void_p()
            _lib.check(self.lib.gi2d_host_pipe_create(C.byref(self._pipe)), "host_pipe_create")
            self._pipe_bufs, self._pipe_idx, self._pipe_dtype = None, 0, None
            self._slot = C.c_int(0)
            self._slot_ref = C.byref(self._slot)
            self._step_host = self.lib.gi2d_fit_step_host
        if self._pipe_bufs is None or self._pipe_dtype != host_img.dtype:
            self._pipe_bufs = [torch.empty(self.H, self.W, 3, dtype=host_img.dtype, device=self.device) for _ in range(2)]
            self._pipe_dtype = host_img.dtype
            self._pipe_bytes = host_img.numel() * host_img.element_size()
            if self.gt_hwc is None:
                self.gt_hwc = self._pipe_bufs[0]
                self.reset_stats(self._step0)
        self.sync_params()
        keep = self.gt_hwc
        bound = []
        for t in self._pipe_bufs:
            self.gt_hwc = t
            self._bind()
            bound.append(C.byref(self.buffers))
            self._pipe_keepalive = getattr(self, "_pipe_keepalive", []) + [self.buffers]
        self._pipe_keepalive = self._pipe_keepalive[-2:]
        self.gt_hwc = keep if keep is not None else self._pipe_bufs[0]
        self._bind()
        self._params_ref = C.byref(self.params)
        self._pipe_bound = bound
        return bound

    def wait_host_result(self, slot: int):
        _lib.check(self.lib.gi2d_host_pipe_wait(self._pipe, int(slot)), "host_pipe_wait")

    def __del__(self):
        pipe = getattr(self, "_pipe", None)
        if pipe is not None and pipe.value:
            try:
                self.lib.gi2d_host_pipe_destroy(pipe)
            except Exception:
                pass

    # ------------------------------------------------------------------ one iteration
    def _enqueue_step(self):
        st = _stream(self.device)
        _lib.check(self.lib.gi2d_fit_forward_backward(C.byref(self.params), C.byref(self.buffers), 1, st),
                   "fit_forward_backward")
        if self.grad_hook is not None:
            self.grad_hook(self)

    def train_iter(self, want_error_map: bool = False):
        """One fit iteration (gaussianimage_covariance.py:249-259), asynchronous.  `want_error_map` also
        leaves the per-pixel L1 error of this iteration's render in `self.err_map` (train.py:87)."""
        if self.gt_hwc is None:
            raise RuntimeError("set_target() first")
        self._dirty = not self.external_optimizer
        self._expected_step += 1
        with torch.cuda.device(self.device):
            self._train_iter_enqueue(want_error_map)
            if getattr(self, "_gt_alt", None) is not None:   # double-buffered targets: this buffer is free after the step
                ev = self._ev_read.get(self.gt_hwc.data_ptr())
                if ev is None:
                    ev = self._ev_read[self.gt_hwc.data_ptr()] = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))

    def _train_iter_enqueue(self, want_error_map: bool):
        if want_error_map:
            self._bind(err_map=True)
            self._enqueue_step()
            self._bind()
            return
        if not self.use_graph or self._eager_left > 0:
            # the first step after (re)allocation runs eagerly: it loads the kernels (CUDA lazy
            # module loading is not capturable) and is an ordinary step in every other respect
            self._eager_left -= 1
            self._enqueue_step()
            return
        if self._graph is None:
            self._bind()
            self._graph = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                with torch.cuda.graph(self._graph, stream=s):
                    self._enqueue_step()
            torch.cuda.current_stream(self.device).wait_stream(s)
        self._graph.replay()

    def train_iters(self, n: int, unroll: int = 8):
        """`n` iterations back to back.  Graph replays of `unroll` steps each: inside a graph the next step's
        first kernel is a programmatic dependent of the rasterizer before it, so its launch latency and
        prologue hide under the previous step's tail (a one-step graph serialises on every replay)."""
        if self.gt_hwc is None:
            raise RuntimeError("set_target() first")
        if not self.use_graph or self.grad_hook is not None or unroll < 2:
            for _ in range(n):
                self.train_iter()
            return
        while n > 0 and self._eager_left > 0:
            self.train_iter()
            n -= 1
        if n >= unroll:
            key = (self.gt_hwc.data_ptr(), unroll)
            g = self._multi_graphs.get(key)
            with torch.cuda.device(self.device):
                if g is None:
                    self._bind()
                    g = torch.cuda.CUDAGraph()
                    s = torch.cuda.Stream(device=self.device)
                    s.wait_stream(torch.cuda.current_stream(self.device))
                    with torch.cuda.stream(s):
                        with torch.cuda.graph(g, stream=s):
                            for _ in range(unroll):
                                self._enqueue_step()
                    torch.cuda.current_stream(self.device).wait_stream(s)
                    self._multi_graphs[key] = g
                for _ in range(n // unroll):
                    g.replay()
            self._expected_step += (n // unroll) * unroll
            self._dirty = not self.external_optimizer
            n %= unroll
        for _ in range(n):
            self.train_iter()

    def launches_per_iter(self, with_backward: bool = True) -> int:
        """Kernels launched by one train_iter / forward (counted by the library itself)."""
        return int(self.lib.gi2d_fit_launch_count(C.byref(self.params), 1 if with_backward else 0))

    # ------------------------------------------------------------------ render / metrics
    def forward(self) -> dict:
        """The model's forward (gaussianimage_covariance.py:187-218): {'render': [1,3,H,W] clamped}.
        A pending optimiser step is applied by the same launch before projecting."""
        rb = self._render_bound
        if rb is None:      # the argument block of a render call (cached: building it costs ~10 us of host time)
            keep = self.buffers
            self._bind(out_img=self.render_chw.data_ptr())
            rb = self._render_bound = (self.buffers, C.byref(self.buffers), C.byref(self.params),
                                       self.render_chw.view(1, 3, self.H, self.W))
            self.buffers = keep
        if torch.cuda.current_device() == self.device.index:
            _lib.check(self.lib.gi2d_fit_forward_backward(rb[2], rb[1], 0, _stream(self.device)),
                       "fit_forward (render)")
        else:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.gi2d_fit_forward_backward(rb[2], rb[1], 0, _stream(self.device)),
                           "fit_forward (render)")
        self._dirty = False
        return {"render": rb[3]}

    __call__ = forward

    def stats(self) -> dict:
        """Synchronises.  mse/psnr refer to the render of the LAST train_iter (before its Adam update),
        like the reference's per-iteration psnr (gaussianimage_covariance.py:256-257)."""
        s = self.stats_buf.cpu()
        self._n, self._n_stale = int(s[STAT_NUM_POINTS]), False   # (the live count rides along for free)
        sse = _sse_total(s)
        mse = sse / (3.0 * self.H * self.W)
        best_mse = float(s[STAT_BEST_SSE]) / (3.0 * self.H * self.W)
        w2, w1, ws = self.loss_w
        loss = w2 * mse
        if w1:
            loss += w1 * float(s[STAT_ABS_SUM]) / (3.0 * self.H * self.W)
        if ws:
            loss += ws * (1.0 - float(s[STAT_SSIM_SUM]) / (3.0 * (self.H - 10) * (self.W - 10)))
        if self.loss_ms[0]:
            loss += self.loss_ms[0] * (1.0 - float(s[STAT_MSSSIM]))
        return {"step": int(s[STAT_STEP]), "num_intersects": int(s[STAT_ISECTS]), "overflow": bool(s[STAT_OVERFLOW]),
                "lr": float(s[STAT_LR]), "sse": sse, "mse": mse, "loss": loss,
                "psnr": 10 * math.log10(1.0 / mse) if mse > 0 else float("inf"),
                "best_sse": float(s[STAT_BEST_SSE]), "best_step": int(s[STAT_BEST_STEP]),
                "best_psnr": (10 * math.log10(1.0 / best_mse) if 0 < best_mse < float("inf") else
                              (0.0 if best_mse > 0 else float("inf"))),
                "non_psd": int(s[STAT_NON_PSD]), "num_points": self._n, "max_tile": int(s[STAT_MAX_TILE])}

    def psnr(self) -> float:
        return self.stats()["psnr"]

    def tile_row_load(self) -> torch.Tensor:
        """Intersections per tile ROW of the last step, f64[tiles_y] on the host: the weights
        `parallel.TileRowPartition(tiles_y, world, row_load=...)` balances the bands with when adaptive
        densification has concentrated the Gaussians (load-balance caveat).  In a band-split run
        every rank sees its own band only: all-reduce (SUM) the result before partitioning.  Synchronises."""
        tx, ty = self.tile_bounds[0], self.tile_bounds[1]
        bins = self.tile_bins
        cnt = (bins[:, 1] - bins[:, 0]).clamp(min=0).view(ty, tx)
        return cnt.sum(dim=1).double().cpu()

    def ms_ssim(self) -> float:
        """MS-SSIM of the current render against the target: the second quality metric the reference reports
        (train.py:190), evaluated by the libgi2d kernels.  Synchronises."""
        from .binding import ms_ssim

        render = self.forward()["render"][0].permute(1, 2, 0).contiguous()
        return ms_ssim(render, self.gt_hwc)

    def stats_async(self, host_slot: torch.Tensor, event: "torch.cuda.Event"):
        """Enqueue a device->host copy of the stats block of the step just issued into pinned `host_slot`
        (f64[STAT_COUNT]) and record `event` after it; the caller waits on the event when it wants the
        numbers.  Lets a driver read EVERY step's result without stalling the GPU between steps."""
        host_slot.copy_(self.stats_buf, non_blocking=True)
        event.record(torch.cuda.current_stream(self.device))

    @staticmethod
    def mse_from_stats(host_slot: torch.Tensor, H: int, W: int) -> float:
        return _sse_total(host_slot) / (3.0 * H * W)

    def ensure_capacity(self, st: Optional[dict] = None) -> bool:
        """Host check of the overflow flag; grows the intersection buffers when it tripped.  Returns True when a
        regrow happened.  A step that overflows is a no-op on the device (no Adam update, step counter / bias
        correction / StepLR untouched: gi2d_fit.cu step_bookkeeping_warp0), so nothing has to be rewound; the
        iterations that did not happen are re-run by `catch_up()`.
        `st`: a stats() result the caller already has (saves the read-back)."""
        st = st if st is not None else self.stats()
        if not st["overflow"]:
            return False
        expected = self._expected_step
        tiles = self.tile_bounds[0] * self.tile_bounds[1]
        # with the bucketed binning EVERY tile's bucket must hold the fullest tile: 1.25 x the largest count when the
        # device reported it, else (coming from the scan + placement path) 4 x the average; at least twice the count
        self._capacity_hint = max(int(st["num_intersects"] * (2 if st.get("max_tile", 0) else 4)),
                                  int(st.get("max_tile", 0) * 1.25 + 8) * tiles)
        self._step0 = st["step"]
        self._alloc_state(zero_moments=False)
        self._reset_keep_best(self._step0, st)
        self._expected_step = expected
        return True

    def catch_up(self, st: Optional[dict] = None) -> dict:
        """Re-run the training iterations that were requested but did not happen because the intersection
        buffers overflowed (each was a no-op on the device): grow the buffers, run `requested - done` more steps,
        until the device step counter has caught up.  Synchronises (one stats read-back per round); `fit()` calls
        it at every host checkpoint, direct users of train_iter()/train_iters() call it before they read results."""
        st = st if st is not None else self.stats()
        while st["step"] < self._expected_step:
            lost = self._expected_step - st["step"]
            if not self.ensure_capacity(st):
                raise _lib.Gi2dError(f"{lost} training steps are missing but no overflow is flagged")
            self._expected_step -= lost
            self.train_iters(lost)
            st = self.stats()
        return st

    def _reset_keep_best(self, step: int, st: dict):
        """reset_stats() that carries the best-so-far squared error / step over (the snapshot stays valid)."""
        keep = self.stats_buf[STAT_SSE:STAT_SSE + STAT_SSE_SLOTS].clone()   # the last step's squared error
        best_n = self.stats_buf[STAT_BEST_N].clone()
        self.reset_stats(step)
        self.stats_buf[STAT_BEST_SSE:STAT_BEST_STEP + 1] = torch.tensor(
            [st["best_sse"], float(st["best_step"])], dtype=torch.float64, device=self.device)
        self.stats_buf[STAT_BEST_N] = best_n
        self.stats_buf[STAT_SSE:STAT_SSE + STAT_SSE_SLOTS] = keep

    def best_state(self) -> dict:
        """The reference's `best_model_dict` + `slv_bound` (train.py:132-137,159-164): parameters right after
        the optimiser step of the best-PSNR iteration, and the SLV bound of that moment -- the device-side
        snapshot (rows [0, best_n) of `best` / `best_bound`; its row count is the model's size at that
        iteration, whatever prune / densify did since).  Synchronises."""
        self.sync_params()
        st = self.stats()
        if st["best_step"] == 0:
            raise RuntimeError("no training step has completed yet")
        n = int(self.stats_buf[STAT_BEST_N].item())
        b = self.best[:n]
        return {"_xyz": b[:, 0:2].clone(), "_cov2d": b[:, 2:5].clone(), "_features_dc": b[:, 5:8].clone(),
                "cholesky_bound": self.best_bound[:n].clone(), "best_step": st["best_step"],
                "best_psnr": st["best_psnr"]}

    def load_best_state(self):
        """train.py:158-164: continue (evaluate) from the best state."""
        bs = self.best_state()
        n = bs["_xyz"].shape[0]
        z = lambda t: torch.zeros_like(t)
        self._replace(bs["_xyz"], bs["_cov2d"], bs["_features_dc"], bs["cholesky_bound"],
                      {"xyz": z(bs["_xyz"]), "cov2d": z(bs["_cov2d"]), "f_dc": z(bs["_features_dc"])},
                      {"xyz": z(bs["_xyz"]), "cov2d": z(bs["_cov2d"]), "f_dc": z(bs["_features_dc"])})
        return n

    # ------------------------------------------------------------------ N-changing operations
    def check_non_semi_definite(self, cov2d=None):
        """gaussianimage_covariance.py:373-382"""
        c = self._cov2d + self.cholesky_bound if cov2d is None else cov2d
        valid = (c[:, 0] * c[:, 2] - c[:, 1] ** 2 > 0) & (c[:, 0] > 0) & (c[:, 2] > 0)
        return int((~valid).sum().item()), valid

    def _set_num_points(self, n: int):
        self._n, self._n_stale = int(n), False
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gi2d_fit_set_num_points(C.byref(self.params), C.byref(self.buffers), int(n),
                                                        _stream(self.device)), "fit_set_num_points")

    def _grow_capacity(self, new_cap: int):
        """More rows than were allocated for: the one case that reallocates (and drops the captured graphs)."""
        self.sync_params()
        n = self.cur_num_points
        st = self.stats()
        expected = self._expected_step

        def grown(t):
            out = torch.zeros(new_cap, t.shape[1], dtype=t.dtype, device=t.device)
            out[:n] = t[:n]
            return out

        self._t_xyz, self._t_cov2d, self._t_f_dc, self._t_bound = (
            grown(t) for t in (self._t_xyz, self._t_cov2d, self._t_f_dc, self._t_bound))
        self._t_m = {k: grown(t) for k, t in self._t_m.items()}
        self._t_v = {k: grown(t) for k, t in self._t_v.items()}
        self.capacity = int(new_cap)
        self._alloc_state(zero_moments=False)
        self._set_num_points(n)
        self._expected_step = expected
        del st

    def _replace(self, xyz, cov, rgb, bound, m, v):
        """Swap in a model of a different size: rows are copied into the capacity-sized arrays (grown first if
        they do not fit), the live count is set, the optimiser's step counter and the best state stay."""
        self.sync_params()
        n = xyz.shape[0]
        if n > self.capacity:
            self._grow_capacity(n)
        for dst, src in ((self._t_xyz, xyz), (self._t_cov2d, cov), (self._t_f_dc, rgb), (self._t_bound, bound)):
            dst[:n].copy_(src)
        for k in _NAMES:
            self._t_m[k][:n].copy_(m[k])
            self._t_v[k][:n].copy_(v[k])
        self._set_num_points(n)

    def prune_async(self):
        """`non_semi_definite_prune` (gaussianimage_covariance.py:354-382) entirely on the device: flush the
        pending step, drop the rows whose covariance + bound is not positive definite (stable compaction of
        parameters, both Adam moments and bounds), update the live count there.  No read-back, no reallocation,
        the step graph stays valid; `cur_num_points` fetches the new count when somebody asks."""
        if self._prune_ws is None:
            nbytes = self.lib.gi2d_fit_prune_workspace_size(self.capacity)
            self._prune_ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
        self._dirty = False
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gi2d_fit_prune(C.byref(self.params), C.byref(self.buffers), self._prune_ws.data_ptr(),
                                               self._prune_ws.numel(), _stream(self.device)), "fit_prune")
        self._n_stale = True

    def non_semi_definite_prune(self):
        """gaussianimage_covariance.py:354-371 with the reference's return value (pruned, remaining): the device
        prune + one read-back of the two counts.  `fit()` uses prune_async() and never reads them."""
        self.prune_async()
        s = self.stats_buf.cpu()
        self._n, self._n_stale = int(s[STAT_NUM_POINTS]), False
        return int(s[STAT_NON_PSD]), self._n

    def densification_postfix(self, new_xyz, new_features_dc, new_cov2d):
        """gaussianimage_covariance.py:307-334: append the given Gaussians (zero Adam moments, new SLV bound).
        (Arbitrary candidates: a torch-side append into the capacity-sized arrays.  The training loop's own
        densification, `add_sample_positions`, selects and appends on the device.)"""
        n_bad, valid = self.check_non_semi_definite(new_cov2d)
        new_xyz, new_features_dc, new_cov2d = new_xyz[valid], new_features_dc[valid], new_cov2d[valid]
        k = new_xyz.shape[0]
        self.sync_params()
        n = self.cur_num_points
        n_new = n + k
        if n_new > self.capacity:
            self._grow_capacity(n_new)
        lp = slv_bound(self.H, self.W, n_new) if self.SLV else 0.5
        self._t_xyz[n:n_new] = new_xyz.to(self._t_xyz)
        self._t_cov2d[n:n_new] = new_cov2d.to(self._t_cov2d)
        self._t_f_dc[n:n_new] = new_features_dc.to(self._t_f_dc)
        self._t_bound[n:n_new] = torch.tensor([lp, 0, lp], dtype=torch.float32, device=self.device)
        for k_ in _NAMES:
            self._t_m[k_][n:n_new] = 0
            self._t_v[k_][n:n_new] = 0
        self._set_num_points(n_new)
        return n_new, n_bad

    def add_sample_positions(self, max_num_points: int, base_num_samples: int = 1000, last: bool = False,
                             errors: Optional[torch.Tensor] = None, new_cov2d: Optional[torch.Tensor] = None):
        """train.py:85-118: new Gaussians at the pixels of largest L1 error, selected AND appended on the device
        (gi2d_fit_densify: a 64-bit radix sort of (error, pixel) keys, the positive-definite filter and the
        append of rows, moments and bounds).  `errors` f32[H,W]: the map a `train_iter(want_error_map=True)` left
        in `self.err_map` (the reference passes that iteration's render); without it the current parameters are
        rendered first.  The candidates' covariances are drawn like the reference does, on the CPU generator
        (train.py:110-112), so the host needs the live count once here: one small read-back per densification.
        `new_cov2d` f32[>=k,3] replaces that draw (parity tests feed the oracle the same numbers)."""
        if errors is None:
            render = self.forward()["render"]
            gt = self.gt_hwc.permute(2, 0, 1).unsqueeze(0)
            if gt.dtype == torch.uint8:
                gt = gt.float() / 255
            errors = torch.abs(render - gt).sum(dim=1)
        if errors.data_ptr() != self.err_map.data_ptr():
            self.err_map.copy_(errors.reshape(self.H, self.W))
        self.sync_params()
        room = max(0, max_num_points - self.cur_num_points)
        k = room if last else min(base_num_samples, room)
        if not k:
            return 0
        if max_num_points > self.capacity:
            self._grow_capacity(max_num_points)
        if new_cov2d is None:
            new_cov2d = torch.rand(k, 3) + torch.tensor([0.5, 0, 0.5])
        cov = new_cov2d[:k].to(self.device, torch.float32).contiguous()
        if self._densify_ws is None:
            nbytes = self.lib.gi2d_fit_densify_workspace_size(self.H, self.W)
            self._densify_ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
        self._bind(err_map=True)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gi2d_fit_densify(C.byref(self.params), C.byref(self.buffers), int(k), cov.data_ptr(),
                                                 1 if self.SLV else 0, self._densify_ws.data_ptr(),
                                                 self._densify_ws.numel(), _stream(self.device)), "fit_densify")
        self._bind()
        self._n_stale = True
        self._keepalive = cov      # (read by the kernel just enqueued)
        return k

    # ------------------------------------------------------------------ checkpoint (train.py:61-77,173-175)
    def save_checkpoint(self, path, best: bool = True, ms_ssim: Optional[float] = None):
        """The reference's `gaussian_model.pth.tar`: {"gs": state_dict, "num_gs", "psnr", "ms-ssim", "slv_bound"},
        with the state-dict keys of GaussianImage_Covariance (`_xyz`, `_cov2d`, `_features_dc` parameters and the
        `_opacity`, `background`, `bound` buffers, gaussianimage_covariance.py:54-70).  `best`: the best-PSNR state
        (what train.py:158-175 saves) instead of the current one.  `ms_ssim`: the value to store; by default the
        MS-SSIM of the CURRENT render (train.py:167,190 evaluate after loading the best state), nan for images
        whose smaller side is <= 160 pixels (pytorch_msssim's own limit)."""
        if ms_ssim is None:
            ms_ssim = self.ms_ssim() if min(self.H, self.W) > 160 and self.gt_hwc is not None else float("nan")
        if best and self.stats()["best_step"] > 0:
            st = self.best_state()
            xyz, cov, rgb, bound, psnr = st["_xyz"], st["_cov2d"], st["_features_dc"], st["cholesky_bound"], st["best_psnr"]
        else:
            xyz, cov, rgb, bound, psnr = self._xyz, self._cov2d, self._features_dc, self.cholesky_bound, self.stats()["psnr"]
        n = xyz.shape[0]
        gs = {"_xyz": xyz.detach().clone(), "_cov2d": cov.detach().clone(), "_features_dc": rgb.detach().clone(),
              "_opacity": torch.ones(n, 1, device=self.device), "background": torch.ones(3, device=self.device),
              "bound": torch.tensor([0.5, 0.5], device=self.device).view(1, 2)}
        torch.save({"gs": gs, "num_gs": n, "psnr": psnr, "ms-ssim": ms_ssim, "slv_bound": bound.detach().clone()}, path)

    def load_checkpoint(self, path):
        """Resume from a reference-format checkpoint (train.py:61-77): parameters and SLV bound are loaded, the
        optimiser state starts from zero (the reference does not store it either)."""
        ck = torch.load(path, map_location=self.device)
        gs = ck["gs"]
        xyz, cov, rgb = (gs[k].to(self.device, torch.float32) for k in ("_xyz", "_cov2d", "_features_dc"))
        bound = ck["slv_bound"].to(self.device, torch.float32)
        if bound.shape[0] != xyz.shape[0]:
            bound = bound.expand(xyz.shape[0], 3)
        z = lambda t: torch.zeros_like(t)
        self._replace(xyz, cov, rgb, bound.contiguous(), {"xyz": z(xyz), "cov2d": z(cov), "f_dc": z(rgb)},
                      {"xyz": z(xyz), "cov2d": z(cov), "f_dc": z(rgb)})
        return ck

    # ------------------------------------------------------------------ the training loop
    def fit(self, iterations: int, max_num_points: Optional[int] = None, prune_iter: int = 100,
            grow_iter: int = 5000, adaptive_add: bool = True, prune: bool = True, callback=None,
            check_iter: int = 250) -> dict:
        """`SimpleTrainer2d.train` (train.py:120-176) without its per-iteration host work: every iteration is one
        graph replay; the best state is snapshotted by the kernels; pruning (every `prune_iter`) compacts the model
        on the device with no read-back; densification (every `grow_iter`) selects and appends on the device after
        ONE small read-back (the live count, for the reference's CPU-generator draw).  The arrays are allocated
        for `max_num_points` up front, so neither changes a pointer nor drops the captured graph.  Every
        `check_iter` iterations (and at the end) the host reads the stats block once to catch an intersection-buffer
        overflow.  Returns the final stats (incl. best_psnr / best_step); `best_state()` has the parameters.
        `callback(iteration, self)` runs after every iteration when given (tests, logging)."""
        max_num_points = max_num_points if max_num_points is not None else self.cur_num_points
        if adaptive_add and max_num_points > self.capacity:
            self._grow_capacity(max_num_points)
        if callback is None:
            # between two host interventions (prune / grow / check / the end) the steps are replayed in unrolled graphs
            it = 0
            while it < iterations:
                nxt = min(iterations, (it // check_iter + 1) * check_iter)
                if prune:
                    nxt = min(nxt, (it // prune_iter + 1) * prune_iter)
                if adaptive_add:
                    nxt = min(nxt, (it // grow_iter + 1) * grow_iter)
                grow = adaptive_add and nxt % grow_iter == 0 and nxt < iterations
                self.train_iters(nxt - it - 1)
                self.train_iter(want_error_map=grow)
                it = nxt
                if prune and it % prune_iter == 0:
                    self.prune_async()
                if grow:
                    self.add_sample_positions(max_num_points, last=(it == iterations - grow_iter), errors=self.err_map)
                if it % check_iter == 0 and it < iterations:
                    st = self.stats()
                    if st["step"] < self._expected_step:
                        self.catch_up(st)
            return self._finish_fit()
        for it in range(1, iterations + 1):
            grow = adaptive_add and it % grow_iter == 0 and it < iterations
            self.train_iter(want_error_map=grow)
            if prune and it % prune_iter == 0:
                self.prune_async()
            if grow:
                self.add_sample_positions(max_num_points, last=(it == iterations - grow_iter), errors=self.err_map)
            if callback is not None:
                callback(it, self)
        return self._finish_fit()

    def _finish_fit(self) -> dict:
        self.sync_params()
        st = self.stats()
        if st["step"] < self._expected_step:   # iterations lost to an intersection-buffer overflow: re-run them
            self.catch_up(st)
            self.sync_params()
            st = self.stats()
        return st
