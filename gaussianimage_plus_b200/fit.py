"""Fused fit loop: the host-side mirror of `GaussianImage_Covariance` for the hot path.

`GaussianImageFitter` keeps the reference model's vocabulary (models/gaussianimage_covariance.py):
`_xyz`, `_cov2d`, `_features_dc`, `cholesky_bound`, `forward()`, `train_iter()`,
`densification_postfix()`, `non_semi_definite_prune()` -- but one `train_iter` is a single CUDA
graph replay of 4 kernels (gi2d_fit.cu) instead of ~70 launches and two host syncs
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
from typing import Optional, Tuple

import torch

from . import _lib
from .binding import TILE, _stream

STAT_STEP, STAT_ISECTS, STAT_OVERFLOW, STAT_LR, STAT_SSE, STAT_SSE_SLOTS = 0, 1, 2, 3, 16, 64
STAT_COUNT = STAT_SSE + STAT_SSE_SLOTS
_NAMES = ("xyz", "cov2d", "f_dc")  # the reference's optimiser group names (gaussianimage_covariance.py:93-96)


def slv_bound(H: int, W: int, num_points: int) -> float:
    """SLV low-pass bound of the reference init: min(HW / (9 pi N), 300) (gaussianimage_covariance.py:61)."""
    return min(H * W / (9 * math.pi * num_points), 300)


def _flushed(attr):
    """Property over a private attribute whose getter first applies a pending optimiser step."""
    def get(self):
        self.sync_params()
        return getattr(self, attr)

    def set_(self, value):
        setattr(self, attr, value)

    return property(get, set_)


class GaussianImageFitter:
    _xyz = _flushed("_t_xyz")                # f32[N,2] pixel coordinates
    _cov2d = _flushed("_t_cov2d")            # f32[N,3] covariance parameters (before the SLV bound)
    _features_dc = _flushed("_t_f_dc")       # f32[N,3] colours
    exp_avg = _flushed("_t_m")               # Adam first moments, dict keyed like the reference's groups
    exp_avg_sq = _flushed("_t_v")            # Adam second moments

    def __init__(self, num_points: int, H: int, W: int, device="cuda:0", lr: float = 0.018,
                 clip_coe: float = 3.0, radius_clip: float = 1.0, color_norm: bool = False,
                 SLV_init: bool = True, tile_rows: Optional[Tuple[int, int]] = None,
                 isect_capacity: Optional[int] = None, use_graph: bool = True,
                 grad_hook=None):
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
        self.grad_hook = grad_hook  # called right after the backward of every step (multi-GPU all-reduce)
        self._capacity_hint = isect_capacity
        self._dirty = False         # a gradient is pending on the device
        self.external_optimizer = False  # True: grad_hook applies the gradients itself (parallel.FusedTileRowExchange)
        self.keep_render = False    # tests: also store the unclamped [H,W,3] render of every train_iter
        f = dict(dtype=torch.float32, device=self.device)
        # reference init, gaussianimage_covariance.py:52-66
        w_init = torch.rand(num_points, 1, **f) * self.W
        h_init = torch.rand(num_points, 1, **f) * self.H
        self._t_xyz = torch.cat((w_init, h_init), dim=1).contiguous()
        self._t_cov2d = torch.rand((num_points, 3), **f)
        self._t_f_dc = torch.zeros(num_points, 3, **f)
        lp = slv_bound(self.H, self.W, num_points) if self.SLV else 0.5
        self.cholesky_bound = torch.tensor([lp, 0, lp], **f).view(1, 3).repeat(num_points, 1).contiguous()
        self.gt_hwc = None
        self._step0 = 0
        self._alloc_state(zero_moments=True)

    # ------------------------------------------------------------------ buffers
    @property
    def cur_num_points(self) -> int:
        return self._t_xyz.shape[0]

    def _raw_params(self):
        return {"xyz": self._t_xyz, "cov2d": self._t_cov2d, "f_dc": self._t_f_dc}

    def _alloc_state(self, zero_moments: bool):
        n = self.cur_num_points
        f = dict(dtype=torch.float32, device=self.device)
        if zero_moments:
            self._t_m = {k: torch.zeros_like(t) for k, t in self._raw_params().items()}
            self._t_v = {k: torch.zeros_like(t) for k, t in self._raw_params().items()}
        tiles = self.tile_bounds[0] * self.tile_bounds[1]
        cap = self._capacity_hint or max(1 << 16, 32 * n)
        self.isect_capacity = int(min(cap, max(n, 1) * tiles, 2 ** 31 - 1024))
        self.grads = torch.zeros(n, 8, **f)
        self.proj = torch.zeros(n, 8, **f)
        self.sorted_keys = torch.zeros(self.isect_capacity, dtype=torch.int64, device=self.device)
        self.tile_bins = torch.zeros(tiles, 2, dtype=torch.int32, device=self.device)
        if not hasattr(self, "stats_buf"):
            self.stats_buf = torch.zeros(STAT_COUNT, dtype=torch.float64, device=self.device)
        self.out_hwc = torch.zeros(self.H, self.W, 3, **f)
        self.render_chw = torch.zeros(3, self.H, self.W, **f)
        self.params = _lib.FitParams(
            n, self.W, self.H, self.tile_bounds[0], self.tile_bounds[1], self.tile_rows[0], self.tile_rows[1],
            self.isect_capacity, self.clip_coe, self.radius_clip, self.lr, 0.9, 0.999, 1e-15, 20000, 0.5,
            int(self.color_norm), 2.0 / (3.0 * self.H * self.W), 1 if self.external_optimizer else 0)
        ws_bytes = self.lib.gi2d_fit_workspace_size(C.byref(self.params))
        self.workspace = torch.zeros(max(ws_bytes, 256), dtype=torch.uint8, device=self.device)
        self._graph = None
        self._eager_left = 1
        self._dirty = False
        self._bind()

    def _bind(self, out_img=None):
        m, v = self._t_m, self._t_v
        if out_img is None and self.keep_render:
            out_img = self.out_hwc.data_ptr()
        gt = self.gt_hwc
        self.buffers = _lib.FitBuffers(
            self._t_xyz.data_ptr(), self._t_cov2d.data_ptr(), self.cholesky_bound.data_ptr(), self._t_f_dc.data_ptr(),
            m["xyz"].data_ptr(), v["xyz"].data_ptr(), m["cov2d"].data_ptr(), v["cov2d"].data_ptr(),
            m["f_dc"].data_ptr(), v["f_dc"].data_ptr(),
            gt.data_ptr() if (gt is not None and gt.dtype == torch.float32) else None,
            out_img, self.grads.data_ptr(), self.proj.data_ptr(), self.sorted_keys.data_ptr(),
            self.tile_bins.data_ptr(), self.stats_buf.data_ptr(), self.workspace.data_ptr(), self.workspace.numel(),
            gt.data_ptr() if (gt is not None and gt.dtype == torch.uint8) else None)

    def reset_stats(self, step: int = 0):
        """Zero the device-side statistics (drops a pending gradient) and set the Adam step counter."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gi2d_fit_reset(C.byref(self.params), C.byref(self.buffers), int(step),
                                               _stream(self.device)), "fit_reset")
        self._dirty = False

    def sync_params(self):
        """Apply a pending optimiser step now (asynchronous; no-op when nothing is pending)."""
        if self._dirty:
            self._dirty = False
            with torch.cuda.device(self.device):
                _lib.check(self.lib.gi2d_fit_adam(C.byref(self.params), C.byref(self.buffers), _stream(self.device)),
                           "fit_adam")

    # ------------------------------------------------------------------ target
    def set_target(self, gt_image: torch.Tensor):
        """gt_image: float32 [1,3,H,W] (the reference's layout, utils.py:21-27) or [H,W,3]; or uint8 [H,W,3]
        -- the image as stored: the kernels then use u8/255, the value ToTensor would have produced, and
        the host->device copy is 4x smaller.  Copied (asynchronously from pinned memory) to HWC on device."""
        if gt_image.dim() == 4:
            gt_image = gt_image[0].permute(1, 2, 0)
        assert gt_image.shape == (self.H, self.W, 3), gt_image.shape
        assert gt_image.dtype in (torch.float32, torch.uint8), gt_image.dtype
        if self.gt_hwc is None or self.gt_hwc.dtype != gt_image.dtype:
            first = self.gt_hwc is None
            self.gt_hwc = torch.empty(self.H, self.W, 3, dtype=gt_image.dtype, device=self.device)
            self._graph = None
            self.gt_hwc.copy_(gt_image, non_blocking=True)
            self._bind()
            if first:
                self.reset_stats(self._step0)
            return
        self.gt_hwc.copy_(gt_image, non_blocking=True)

    # ------------------------------------------------------------------ one iteration
    def _enqueue_step(self):
        st = _stream(self.device)
        _lib.check(self.lib.gi2d_fit_forward_backward(C.byref(self.params), C.byref(self.buffers), 1, st),
                   "fit_forward_backward")
        if self.grad_hook is not None:
            self.grad_hook(self)

    def train_iter(self):
        """One fit iteration (gaussianimage_covariance.py:249-259), asynchronous."""
        if self.gt_hwc is None:
            raise RuntimeError("set_target() first")
        self._dirty = not self.external_optimizer
        with torch.cuda.device(self.device):
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

    def launches_per_iter(self, with_backward: bool = True) -> int:
        """Kernels launched by one train_iter / forward (counted by the library itself)."""
        return int(self.lib.gi2d_fit_launch_count(C.byref(self.params), 1 if with_backward else 0))

    # ------------------------------------------------------------------ render / metrics
    def forward(self) -> dict:
        """The model's forward (gaussianimage_covariance.py:187-218): {'render': [1,3,H,W] clamped}.
        A pending optimiser step is applied by the same launch before projecting."""
        with torch.cuda.device(self.device):
            self._bind(out_img=self.render_chw.data_ptr())
            _lib.check(self.lib.gi2d_fit_forward_backward(C.byref(self.params), C.byref(self.buffers), 0,
                                                          _stream(self.device)), "fit_forward (render)")
            self._bind(out_img=None)
        self._dirty = False
        return {"render": self.render_chw.view(1, 3, self.H, self.W)}

    __call__ = forward

    def stats(self) -> dict:
        """Synchronises.  mse/psnr refer to the render of the LAST train_iter (before its Adam update),
        like the reference's per-iteration psnr (gaussianimage_covariance.py:256-257)."""
        s = self.stats_buf.cpu()
        sse = float(s[STAT_SSE:STAT_SSE + STAT_SSE_SLOTS].sum())
        mse = sse / (3.0 * self.H * self.W)
        return {"step": int(s[STAT_STEP]), "num_intersects": int(s[STAT_ISECTS]), "overflow": bool(s[STAT_OVERFLOW]),
                "lr": float(s[STAT_LR]), "sse": sse, "mse": mse,
                "psnr": 10 * math.log10(1.0 / mse) if mse > 0 else float("inf")}

    def psnr(self) -> float:
        return self.stats()["psnr"]

    def stats_async(self, host_slot: torch.Tensor, event: "torch.cuda.Event"):
        """Enqueue a device->host copy of the stats block of the step just issued into pinned `host_slot`
        (f64[STAT_COUNT]) and record `event` after it; the caller waits on the event when it wants the
        numbers.  Lets a driver read EVERY step's result without stalling the GPU between steps."""
        host_slot.copy_(self.stats_buf, non_blocking=True)
        event.record(torch.cuda.current_stream(self.device))

    @staticmethod
    def mse_from_stats(host_slot: torch.Tensor, H: int, W: int) -> float:
        return float(host_slot[STAT_SSE:STAT_SSE + STAT_SSE_SLOTS].sum()) / (3.0 * H * W)

    def ensure_capacity(self) -> bool:
        """Host check of the overflow flag; grows the intersection buffers when it tripped.
        Returns True when a regrow happened (the overflowing step's gradient is dropped, not applied)."""
        st = self.stats()
        if not st["overflow"]:
            return False
        self._capacity_hint = int(st["num_intersects"] * 2)
        self._step0 = st["step"] - 1  # the overflowing step did not update the parameters
        self._alloc_state(zero_moments=False)
        self.reset_stats(self._step0)
        return True

    # ------------------------------------------------------------------ N-changing operations
    def check_non_semi_definite(self, cov2d=None):
        """gaussianimage_covariance.py:373-382"""
        c = self._cov2d + self.cholesky_bound if cov2d is None else cov2d
        valid = (c[:, 0] * c[:, 2] - c[:, 1] ** 2 > 0) & (c[:, 0] > 0) & (c[:, 2] > 0)
        return int((~valid).sum().item()), valid

    def _replace(self, xyz, cov, rgb, bound, m, v):
        step = self.stats()["step"]
        self._t_xyz, self._t_cov2d, self._t_f_dc, self.cholesky_bound = (
            t.contiguous() for t in (xyz, cov, rgb, bound))
        self._t_m, self._t_v = m, v
        self._step0 = step
        self._alloc_state(zero_moments=False)
        self.reset_stats(step)

    def non_semi_definite_prune(self):
        """gaussianimage_covariance.py:354-371: drop Gaussians whose covariance is not positive definite
        (parameters, Adam moments and SLV bounds are masked together)."""
        n_bad, valid = self.check_non_semi_definite()      # (flushes a pending step first)
        if n_bad and self.cur_num_points - n_bad > 0:
            m = {k: t[valid].contiguous() for k, t in self.exp_avg.items()}
            v = {k: t[valid].contiguous() for k, t in self.exp_avg_sq.items()}
            self._replace(self._xyz[valid], self._cov2d[valid], self._features_dc[valid],
                          self.cholesky_bound[valid], m, v)
        return n_bad, self.cur_num_points

    def densification_postfix(self, new_xyz, new_features_dc, new_cov2d):
        """gaussianimage_covariance.py:307-334: append Gaussians (zero Adam moments, new SLV bound)."""
        n_bad, valid = self.check_non_semi_definite(new_cov2d)
        new_xyz, new_features_dc, new_cov2d = new_xyz[valid], new_features_dc[valid], new_cov2d[valid]
        k = new_xyz.shape[0]
        cat = lambda a, b: torch.cat((a, b.to(a)), dim=0)
        n_new = self.cur_num_points + k
        lp = slv_bound(self.H, self.W, n_new) if self.SLV else 0.5
        new_bound = torch.tensor([lp, 0, lp], dtype=torch.float32, device=self.device).view(1, 3).repeat(k, 1)
        names = dict(xyz=new_xyz, cov2d=new_cov2d, f_dc=new_features_dc)
        m = {kk: cat(t, torch.zeros_like(names[kk])) for kk, t in self.exp_avg.items()}
        v = {kk: cat(t, torch.zeros_like(names[kk])) for kk, t in self.exp_avg_sq.items()}
        self._replace(cat(self._xyz, new_xyz), cat(self._cov2d, new_cov2d), cat(self._features_dc, new_features_dc),
                      cat(self.cholesky_bound, new_bound), m, v)
        return n_new, n_bad

    def add_sample_positions(self, max_num_points: int, base_num_samples: int = 1000, last: bool = False):
        """train.py:85-118: new Gaussians at the pixels of largest L1 error of the current render."""
        render = self.forward()["render"]
        gt = self.gt_hwc.permute(2, 0, 1).unsqueeze(0)
        if gt.dtype == torch.uint8:
            gt = gt.float() / 255
        errors = torch.abs(render - gt).sum(dim=1)
        p_flat = (errors / torch.sum(errors)).view(-1)
        room = max(0, max_num_points - self.cur_num_points)
        k = room if last else min(base_num_samples, room)
        if not k:
            return 0
        _, idx = torch.topk(p_flat, k)
        xyz = torch.stack([idx % self.W, idx // self.W], dim=1).float()
        color = torch.zeros(k, 3, device=self.device)
        cov = torch.rand(k, 3, device=self.device) + torch.tensor([0.5, 0, 0.5], device=self.device)
        self.densification_postfix(xyz, color, cov)
        return k
