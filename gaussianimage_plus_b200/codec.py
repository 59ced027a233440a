"""Compression pass of GaussianImage++ (`train_quantize.py`, BASELINE.json configs[3]): the
`quantize=True` half of `GaussianImage_Covariance` -- `training_setup(quantize=True)`
(models/gaussianimage_covariance.py:116-146), `forward_quantize` (:384-410), `train_iter_quantize` (:219-232),
`optimizer_step` (:234-247), `compress_wo_ec` / `decompress_wo_ec` / `analysis_wo_ec` (:412-509) -- on top of
the drop-in operators (`gsplat.project_gaussians_2d_covariance`, `gsplat.rasterize_gaussians_plus`: the CUDA
kernels of libgi2d with the reference's autograd contract) and the fused quantisers of `quantize.py`.

Two forms of the quantisation-aware iteration:

* `QuantizedGaussianImage` -- the reference's structure, operator by operator: quantisers -> projection ->
  rasterization (the libgi2d operators with the reference's autograd contract) -> `loss_fn` (the fused SSIM /
  mse / l1 kernels) -> four optimisers; one host read-back per iteration like the reference; ~3 ms/iteration.
* `FusedQuantizedTrainer` -- same quantisers, same optimisers, but everything between the de-quantised
  attributes and their gradients is the fused fit step (3 kernels, `external_optimizer` mode +
  gi2d_fit_input_grads), nothing synchronises, and the whole iteration is replayed from ONE CUDA graph:
  ~0.5 ms/iteration.

The fully fused, graph-captured step of `fit.py` is the warm-up phase (`iter < warmup_iter`,
train_quantize.py:124-127).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
from torch import nn

from .fit import loss_weights, msssim_term
from .quantize import HybirdQuant, UniformQuantizer


class _ImageLoss(torch.autograd.Function):
    """loss_fn of models/utils.py:60-80 on a [1,3,H,W] prediction in [0,1], value and gradient from the libgi2d
    loss kernels (gi2d_image_loss_grad: SSIM map + its transposed filter, mse / l1 terms)."""

    @staticmethod
    def forward(ctx, pred, target_hwc, loss_type, lambda_value):
        from .binding import image_loss_grad

        hwc = pred[0].permute(1, 2, 0).contiguous()
        v, ssim_sum = image_loss_grad(hwc, target_hwc, loss_type, lambda_value)
        H, W, _ = hwc.shape
        w2, w1, ws = loss_weights(loss_type, lambda_value)
        tgt = target_hwc.float() / 255 if target_hwc.dtype == torch.uint8 else target_hwc
        d = hwc.clamp(0, 1) - tgt
        loss = w2 * (d * d).mean() + w1 * d.abs().mean()
        if ws:
            loss = loss + ws * (1.0 - ssim_sum[0].float() / (3.0 * (H - 10) * (W - 10)))
        wm, _ = msssim_term(loss_type, lambda_value)
        if wm:                                      # (image_loss_grad hands back ms_ssim itself for these)
            loss = loss + wm * (1.0 - ssim_sum[0].float())
        ctx.save_for_backward(v)
        return loss

    @staticmethod
    def backward(ctx, g):
        (v,) = ctx.saved_tensors
        return (g * v).permute(2, 0, 1).unsqueeze(0), None, None, None


def loss_fn(pred: torch.Tensor, target_hwc: torch.Tensor, loss_type: str = "L2", lambda_value: float = 0.7):
    """pred [1,3,H,W]; target f32 or u8 [H,W,3] on the same device."""
    return _ImageLoss.apply(pred, target_hwc, loss_type, lambda_value)


class QuantizedGaussianImage(nn.Module):
    """State + quantisers + the four optimisers of the reference's quantised model."""

    def __init__(self, xyz, cov2d, features_dc, cholesky_bound, H: int, W: int, lr: float = 0.018,
                 loss_type: str = "L2", color_norm: bool = False, xy_bit: int = 12, cov_bit: int = 10,
                 color_bit: int = 6, clip_coe: float = 3.0, radius_clip: float = 1.0):
        super().__init__()
        self.H, self.W = int(H), int(W)
        self.tile_bounds = ((self.W + 15) // 16, (self.H + 15) // 16, 1)
        self.loss_type, self.color_norm = loss_type, bool(color_norm)
        self.gs_clip_coe, self.radius_clip = clip_coe, radius_clip
        self._xyz = nn.Parameter(xyz.detach().clone())
        self._cov2d = nn.Parameter(cov2d.detach().clone())
        self._features_dc = nn.Parameter(features_dc.detach().clone())
        self.register_buffer("_opacity", torch.ones(xyz.shape[0], 1, device=xyz.device))
        self.cholesky_bound = cholesky_bound.detach().clone()
        self.cur_num_points = xyz.shape[0]
        dev = xyz.device
        # training_setup(lr, update_optimizer=True, quantize=True), gaussianimage_covariance.py:105-146
        groups = [{"params": [self._xyz], "lr": lr, "name": "xyz"},
                  {"params": [self._features_dc], "lr": lr, "name": "f_dc"},
                  {"params": [self._cov2d], "lr": lr, "name": "cov2d"}]
        self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=20000, gamma=0.5)
        qlr = 0.001
        self.xyz_quantizer = UniformQuantizer(signed=False, bits=xy_bit, weight=1.0, learned=True, num_channels=2).to(dev)
        self.xyz_quantizer_optimizer = torch.optim.Adam(self.xyz_quantizer.parameters(), lr=qlr)
        self.xyz_scheduler = torch.optim.lr_scheduler.StepLR(self.xyz_quantizer_optimizer, step_size=10000, gamma=0.5)
        self.cholesky_quantizer = HybirdQuant(signed=False, bits=cov_bit, cov_bits=cov_bit, learned=True, weight=1.0).to(dev)
        self.cov2d_quantizer_optimizer = torch.optim.Adam(self.cholesky_quantizer.parameters(), lr=qlr, eps=1e-15)
        self.cov2d_scheduler = torch.optim.lr_scheduler.StepLR(self.cov2d_quantizer_optimizer, step_size=10000, gamma=0.5)
        self.features_dc_quantizer = UniformQuantizer(signed=False, bits=color_bit, learned=True, weight=1.0,
                                                      num_channels=3).to(dev)
        self.color_quantizer_optimizer = torch.optim.Adam(self.features_dc_quantizer.parameters(), lr=qlr, eps=1e-15)
        self.color_scheduler = torch.optim.lr_scheduler.StepLR(self.color_quantizer_optimizer, step_size=10000, gamma=0.5)
        self.quantized_cov2d = None

    @classmethod
    def from_fitter(cls, fit, best: bool = True, **kw):
        """Continue from a fitted model (train_quantize.py:128-140 restarts from the best warm-up state)."""
        if best:
            st = fit.best_state()
            args = (st["_xyz"], st["_cov2d"], st["_features_dc"], st["cholesky_bound"])
        else:
            args = (fit._xyz, fit._cov2d, fit._features_dc, fit.cholesky_bound)
        kw.setdefault("lr", float(fit.stats()["lr"]))       # scheduler.get_last_lr() of the warm-up (:137)
        kw.setdefault("loss_type", fit.loss_type)
        kw.setdefault("color_norm", fit.color_norm)
        return cls(*args, fit.H, fit.W, clip_coe=fit.clip_coe, radius_clip=fit.radius_clip, **kw)

    # ------------------------------------------------------------------ reference properties
    @property
    def get_cov2d_elements(self):
        return self._cov2d + self.cholesky_bound

    @property
    def get_features(self):
        return torch.sigmoid(self._features_dc) if self.color_norm else self._features_dc

    @property
    def get_opacity(self):
        return self._opacity

    def _render(self, means, cov, colors):
        from .gsplat import project_gaussians_2d_covariance, rasterize_gaussians_plus

        xys, depths, radii, conics, nth = project_gaussians_2d_covariance(
            means, cov, self.H, self.W, self.tile_bounds, clip_coe=self.gs_clip_coe, radius_clip=self.radius_clip)
        out = rasterize_gaussians_plus(xys, depths, radii, conics, nth, colors, self.get_opacity, self.H, self.W, 16, 16,
                                       radius_clip=self.radius_clip)
        out = torch.clamp(out, 0, 1)
        return out.view(-1, self.H, self.W, 3).permute(0, 3, 1, 2).contiguous()

    # ------------------------------------------------------------------ gaussianimage_covariance.py:384-410
    def forward_quantize(self) -> Dict:
        means, l_vqm, m_bit, _ = self.xyz_quantizer(self._xyz)
        cov, l_vqs, s_bit, _ = self.cholesky_quantizer(self.get_cov2d_elements)
        self.quantized_cov2d = cov
        colors, l_vqc, c_bit, _ = self.features_dc_quantizer(self.get_features)
        return {"render": self._render(means, cov, colors), "vq_loss": l_vqm + l_vqs + l_vqc,
                "unit_bit": [m_bit, s_bit, c_bit]}

    # ------------------------------------------------------------------ :219-247
    def train_iter_quantize(self, gt_hwc: torch.Tensor):
        """gt_hwc: f32 or u8 [H,W,3].  Returns (image, loss, img_loss, vq_loss, psnr) like the reference."""
        pkg = self.forward_quantize()
        image = pkg["render"]
        loss = loss_fn(image, gt_hwc, self.loss_type, 0.7)
        loss.backward()
        self.optimizer_step()
        with torch.no_grad():
            tgt = gt_hwc.float() / 255 if gt_hwc.dtype == torch.uint8 else gt_hwc
            mse = ((image[0].permute(1, 2, 0) - tgt) ** 2).mean().item()
            psnr = 10 * math.log10(1.0 / mse)
        return image.detach(), loss, float(loss.detach()), pkg["vq_loss"], psnr

    def optimizer_step(self):
        for opt, sch in ((self.optimizer, self.scheduler), (self.cov2d_quantizer_optimizer, self.cov2d_scheduler),
                         (self.xyz_quantizer_optimizer, self.xyz_scheduler),
                         (self.color_quantizer_optimizer, self.color_scheduler)):
            opt.step()
            opt.zero_grad(set_to_none=True)
            sch.step()

    # ------------------------------------------------------------------ :373-382, :412-467
    def check_non_semi_definite(self, cov2d):
        valid = (cov2d[:, 0] * cov2d[:, 2] - cov2d[:, 1] ** 2 > 0) & (cov2d[:, 0] > 0) & (cov2d[:, 2] > 0)
        return int((~valid).sum().item()), valid

    @torch.no_grad()
    def compress_wo_ec(self) -> Dict:
        means, quant_means = self.xyz_quantizer.compress(self._xyz)
        cov, quant_cov = self.cholesky_quantizer.compress(self.get_cov2d_elements)
        colors, color_index = self.features_dc_quantizer.compress(self.get_features)
        n_bad, valid = self.check_non_semi_definite(cov)
        if n_bad:       # Gaussians that stopped being positive definite after quantisation are dropped (:424-437)
            cov, quant_cov, means = cov[valid], quant_cov[valid], means[valid]
            color_index, colors = color_index[valid], colors[valid]
            quant_means = quant_means[valid]
            self.cholesky_bound = self.cholesky_bound[valid]
            self._opacity = torch.ones(int(valid.sum()), 1, device=means.device)
            self.cur_num_points = int(valid.sum())
        self.quantized_cov2d = cov
        return {"xyz": means, "feature_dc_index": color_index, "quant_cholesky_elements": quant_cov,
                "quant_means": quant_means}

    @torch.no_grad()
    def decompress_wo_ec(self, enc: Dict) -> Dict:
        cov = self.cholesky_quantizer.decompress(enc["quant_cholesky_elements"])
        colors = self.features_dc_quantizer.decompress(enc["feature_dc_index"])
        return {"render": self._render(enc["xyz"].contiguous(), cov.contiguous(), colors.contiguous())}

    def analysis_wo_ec(self, enc: Dict) -> Dict:
        """Bits per pixel without entropy coding (:469-509, `lsq` branches): N x bits per attribute + the
        quantiser parameters (32-bit scale and beta per channel)."""
        cov_bits = enc["quant_cholesky_elements"].numel() * self.cholesky_quantizer.size() + 32 * 3 * 2
        color_bits = enc["feature_dc_index"].numel() * self.features_dc_quantizer.size() + 32 * 3 * 2
        pos_bits = enc["xyz"].numel() * self.xyz_quantizer.size() + 32 * 2 * 2
        px = self.H * self.W
        return {"bpp": (pos_bits + cov_bits + color_bits) / px, "position_bpp": pos_bits / px,
                "cholesky_bpp": cov_bits / px, "feature_dc_bpp": color_bits / px}


# ----------------------------------------------------------------------------------- fused QAT
class _FusedRenderLoss(torch.autograd.Function):
    """loss(means, cov, colors): projection + binning + rasterize forward + loss + rasterize backward +
    projection backward are the 3 kernels of the fused fit step (gi2d_fit.cu, `external_optimizer` mode: the
    step leaves parameters and optimiser to the caller) + gi2d_fit_input_grads; no host synchronisation, so the
    whole quantisation-aware iteration can live in one CUDA graph."""

    @staticmethod
    def forward(ctx, means, cov, colors, owner):
        import ctypes as C

        from . import _lib
        from .binding import _stream

        fit = owner._fit
        fit._t_xyz.copy_(means)
        fit._t_cov2d.copy_(cov)
        fit._t_f_dc.copy_(colors)
        fit._enqueue_step()
        _lib.check(fit.lib.gi2d_fit_input_grads(C.byref(fit.params), C.byref(fit.buffers), owner._gbuf.data_ptr(),
                                                _stream(fit.device)), "fit_input_grads")
        ctx.owner = owner
        return owner._loss_from_stats()

    @staticmethod
    def backward(ctx, g):
        gb = ctx.owner._gbuf
        return g * gb[:, 0:2], g * gb[:, 2:5], g * gb[:, 5:8], None


class FusedQuantizedTrainer(QuantizedGaussianImage):
    """`QuantizedGaussianImage` whose iteration runs sync-free: the quantisers and the four optimisers stay the
    torch modules / torch.optim.Adam of the parent class (capturable), everything between the de-quantised
    attributes and their gradients is the fused fit step, and forward + backward + optimiser steps are replayed
    from ONE CUDA graph (re-captured when a StepLR changes a learning rate).  PSNR is read on demand
    (`psnr()`), not every iteration."""

    def __init__(self, *a, use_graph: bool = True, **kw):
        super().__init__(*a, **kw)
        from .fit import GaussianImageFitter

        dev = self._xyz.device
        n = self._xyz.shape[0]
        fit = GaussianImageFitter(n, self.H, self.W, device=dev, lr=0.0, clip_coe=self.gs_clip_coe,
                                  radius_clip=self.radius_clip, color_norm=False, use_graph=False,
                                  loss_type=self.loss_type)
        fit.external_optimizer = True          # the step never touches parameters / moments
        fit.track_best = False
        fit.params.external_optimizer = 1
        fit.cholesky_bound.zero_()             # the inputs are complete covariances (bound already added)
        fit._bind()
        self._fit = fit
        self._gbuf = torch.zeros(n, 8, device=dev)
        self.use_graph = use_graph
        self._graph = None
        self._graph_lrs = None
        self._side = None
        self._warm = 0
        for opt in (self.optimizer, self.cov2d_quantizer_optimizer, self.xyz_quantizer_optimizer,
                    self.color_quantizer_optimizer):
            for gr in opt.param_groups:
                gr["capturable"] = True

    def set_target(self, gt_hwc: torch.Tensor):
        self._fit.set_target(gt_hwc)

    def _loss_from_stats(self) -> torch.Tensor:
        from .fit import STAT_ABS_SUM, STAT_SSE, STAT_SSE_SLOTS, STAT_SSIM_SUM

        s, f = self._fit.stats_buf, self._fit
        w2, w1, ws = f.loss_w
        px3 = 3.0 * self.H * self.W
        loss = w2 * s[STAT_SSE:STAT_SSE + STAT_SSE_SLOTS].sum() / px3
        if w1:
            loss = loss + w1 * s[STAT_ABS_SUM] / px3
        if ws:
            loss = loss + ws * (1.0 - s[STAT_SSIM_SUM] / (3.0 * (self.H - 10) * (self.W - 10)))
        if f.loss_ms[0]:
            from .fit import STAT_MSSSIM

            loss = loss + f.loss_ms[0] * (1.0 - s[STAT_MSSSIM])
        return loss.float()

    def _iteration(self):
        means, _, _, _ = self.xyz_quantizer(self._xyz)
        cov, _, _, _ = self.cholesky_quantizer(self.get_cov2d_elements)
        colors, _, _, _ = self.features_dc_quantizer(self.get_features)
        loss = _FusedRenderLoss.apply(means, cov, colors, self)
        loss.backward()
        for opt in (self.optimizer, self.cov2d_quantizer_optimizer, self.xyz_quantizer_optimizer,
                    self.color_quantizer_optimizer):
            opt.step()
        return loss

    def _lrs(self):
        return tuple(g["lr"] for o in (self.optimizer, self.cov2d_quantizer_optimizer, self.xyz_quantizer_optimizer,
                                       self.color_quantizer_optimizer) for g in o.param_groups)

    def train_iter_quantize(self, gt_hwc: Optional[torch.Tensor] = None):
        """One quantisation-aware iteration (gaussianimage_covariance.py:219-247), asynchronous; returns the
        loss as a device scalar.  `gt_hwc` (optional) replaces the target first."""
        if gt_hwc is not None:
            self._fit.set_target(gt_hwc)
        opts = (self.optimizer, self.cov2d_quantizer_optimizer, self.xyz_quantizer_optimizer,
                self.color_quantizer_optimizer)
        scheds = (self.scheduler, self.cov2d_scheduler, self.xyz_scheduler, self.color_scheduler)
        dev = self._xyz.device
        cur = torch.cuda.current_stream(dev)
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        side = self._side
        if not self.use_graph or self._warm < 3:
            # eager: the first call initialises the quantisers from the data (quantize.py:72-80), the optimisers
            # create their state, the kernels get loaded -- none of which can be captured.  It runs on the SAME
            # side stream the capture will use: autograd ties the parameters' gradient accumulation to the stream
            # of the first backward, and a capture may not depend on the legacy default stream.
            self._warm += 1
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for o in opts:
                    o.zero_grad(set_to_none=True)
                loss = self._iteration().detach()
            cur.wait_stream(side)
        else:
            if self._graph is None or self._graph_lrs != self._lrs():
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    for o in opts:
                        o.zero_grad(set_to_none=True)
                    self._graph = torch.cuda.CUDAGraph()
                    self._graph_lrs = self._lrs()
                    with torch.cuda.graph(self._graph, stream=side):
                        self._static_loss = self._iteration().detach()
                cur.wait_stream(side)
            self._graph.replay()
            loss = self._static_loss
        for sc in scheds:
            sc.step()
        return loss

    def psnr(self) -> float:
        """PSNR of the render of the LAST iteration (before its update); synchronises."""
        return self._fit.stats()["psnr"]


# ----------------------------------------------------------------------------------- QAT as kernels
class KernelQuantizedTrainer(QuantizedGaussianImage):
    """The quantisation-aware iteration with the quantisers, their straight-through backward and the four
    optimisers as KERNELS (csrc/gi2d_quant.cu) around the fused fit step:

        gi2d_quant_forward (2 launches)  ->  fit step, external_optimizer (3)  ->  gi2d_fit_input_grads (1)
        ->  gi2d_quant_backward_step (2)

    eight launches replayed from one CUDA graph, no torch autograd, no torch.optim; the learning-rate schedules
    are evaluated on the device from the iteration counter, so the graph is captured once.  The attributes are
    the parent class's nn.Parameters (updated in place), the 12 quantiser parameters live in `qparams` and are
    copied into the parent's quantiser modules by `sync_quantizers()` (compress_wo_ec does it for you)."""

    def __init__(self, *a, use_graph: bool = True, debug_grads: bool = False, **kw):
        super().__init__(*a, **kw)
        import ctypes as C

        from . import _lib
        from .fit import GaussianImageFitter

        dev = self._xyz.device
        n = self._xyz.shape[0]
        fit = GaussianImageFitter(n, self.H, self.W, device=dev, lr=0.0, clip_coe=self.gs_clip_coe,
                                  radius_clip=self.radius_clip, color_norm=False, use_graph=False,
                                  loss_type=self.loss_type)
        fit.external_optimizer = True
        fit.track_best = False
        fit.params.external_optimizer = 1
        fit.cholesky_bound.zero_()             # the inputs are complete covariances (bound already added)
        fit._bind()
        self._fit = fit
        f = dict(device=dev, dtype=torch.float32)
        self._gbuf = torch.zeros(n, 8, **f)
        self._moments = {k: torch.zeros_like(p) for k, p in
                         (("m_xyz", self._xyz), ("v_xyz", self._xyz), ("m_cov", self._cov2d), ("v_cov", self._cov2d),
                          ("m_rgb", self._features_dc), ("v_rgb", self._features_dc))}
        self.qparams = torch.zeros(12, **f)
        self._qm, self._qv = torch.zeros(12, **f), torch.zeros(12, **f)
        self.qstats = torch.zeros(32, device=dev, dtype=torch.float64)
        self.dbg_grads = torch.zeros(n, 8, **f) if debug_grads else None
        lr = float(self.optimizer.param_groups[0]["lr"])
        self._qp = _lib.QuantParams(n, self.xyz_quantizer.qmax, self.cholesky_quantizer.cov_quantizer.qmax,
                                    self.features_dc_quantizer.qmax, int(self.color_norm), lr, 20000, 0.001, 10000,
                                    0.5, 0.9, 0.999, 1e-15, 1e-8)
        if self.cholesky_quantizer.var_quantizer.qmax != self.cholesky_quantizer.cov_quantizer.qmax:
            raise NotImplementedError("the kernels assume bits == cov_bits in HybirdQuant (the reference's setting)")
        bound = self.cholesky_bound.contiguous()
        self.cholesky_bound = bound
        ptr = lambda t: t.data_ptr() if t is not None else None
        self._qb = _lib.QuantBuffers(
            ptr(self._xyz.data), ptr(self._cov2d.data), ptr(self._features_dc.data), ptr(bound),
            *(ptr(self._moments[k]) for k in ("m_xyz", "v_xyz", "m_cov", "v_cov", "m_rgb", "v_rgb")),
            ptr(self.qparams), ptr(self._qm), ptr(self._qv), ptr(self.qstats),
            ptr(fit._t_xyz), ptr(fit._t_cov2d), ptr(fit._t_f_dc), ptr(self._gbuf), ptr(self.dbg_grads))
        self._C = C
        self.use_graph = use_graph
        self._graph = None
        self._warm = 0
        self._initialised = False
        self.iterations = 0

    def set_target(self, gt_hwc: torch.Tensor):
        self._fit.set_target(gt_hwc)

    def _enqueue_iteration(self):
        from . import _lib
        from .binding import _stream

        C, fit = self._C, self._fit
        st = _stream(fit.device)
        _lib.check(fit.lib.gi2d_quant_forward(C.byref(self._qp), C.byref(self._qb), st), "quant_forward")
        fit._enqueue_step()
        _lib.check(fit.lib.gi2d_fit_input_grads(C.byref(fit.params), C.byref(fit.buffers), self._gbuf.data_ptr(), st),
                   "fit_input_grads")
        _lib.check(fit.lib.gi2d_quant_backward_step(C.byref(self._qp), C.byref(self._qb), st), "quant_backward_step")

    def init_quantizers(self):
        """`_init_data` of the three quantisers from the current attributes (first call of the reference's
        forward_quantize, quantize.py:72-80 / 352-354)."""
        from . import _lib
        from .binding import _stream

        _lib.check(self._fit.lib.gi2d_quant_init(self._C.byref(self._qp), self._C.byref(self._qb),
                                                 _stream(self._fit.device)), "quant_init")
        self._initialised = True

    def train_iter_quantize(self, gt_hwc: Optional[torch.Tensor] = None):
        """One quantisation-aware iteration (gaussianimage_covariance.py:219-247), asynchronous.  Returns
        nothing: read `psnr()` / `loss()` when needed."""
        if gt_hwc is not None:
            self._fit.set_target(gt_hwc)
        dev = self._xyz.device
        with torch.cuda.device(dev):
            if not self._initialised:
                self.init_quantizers()
            if not self.use_graph or self._warm < 2:
                self._warm += 1
                self._enqueue_iteration()
            else:
                if self._graph is None:
                    self._graph = torch.cuda.CUDAGraph()
                    s = torch.cuda.Stream(device=dev)
                    s.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(s):
                        with torch.cuda.graph(self._graph, stream=s):
                            self._enqueue_iteration()
                    torch.cuda.current_stream(dev).wait_stream(s)
                self._graph.replay()
        self.iterations += 1

    def psnr(self) -> float:
        """PSNR of the render of the LAST iteration (before its update); synchronises.  Also the place where an
        overflow of the intersection buffers is noticed (the iteration then applied zero gradients)."""
        st = self._fit.stats()
        if st["overflow"]:
            raise RuntimeError("intersection buffers overflowed during quantisation-aware training: construct the "
                               "trainer's fitter with a larger capacity")
        return st["psnr"]

    def loss(self) -> float:
        return float(self._loss_from_stats())

    _loss_from_stats = FusedQuantizedTrainer._loss_from_stats

    @torch.no_grad()
    def sync_quantizers(self):
        """qparams -> the scale / beta parameters of the parent's quantiser modules (state-dict compatible)."""
        if not self._initialised:
            self.init_quantizers()
        q = self.qparams
        self.xyz_quantizer.scale.data.copy_(q[0:2])
        self.xyz_quantizer.beta.data.copy_(q[2:4])
        self.cholesky_quantizer.cov_quantizer.scale.data.copy_(q[4:5])
        self.cholesky_quantizer.cov_quantizer.beta.data.copy_(q[5:6])
        self.features_dc_quantizer.scale.data.copy_(q[6:9])
        self.features_dc_quantizer.beta.data.copy_(q[9:12])
        for m in (self.xyz_quantizer, self.cholesky_quantizer, self.cholesky_quantizer.cov_quantizer,
                  self.cholesky_quantizer.var_quantizer, self.features_dc_quantizer):
            m.init_state = 1

    def forward_quantize(self) -> Dict:
        self.sync_quantizers()
        return super().forward_quantize()

    def compress_wo_ec(self) -> Dict:
        self.sync_quantizers()
        return super().compress_wo_ec()
