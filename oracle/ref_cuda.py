"""TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference CUDA extension built by
oracle/build_ref.py (oracle/_ref/<variant>/gsplat_ref_<variant>.so) on the GPU box.

The reference's Python wrappers cannot travel (they live under /root/reference), so this module
restates their *call order* only -- which `_C` binding is called with which arguments, and the
three torch ops between them (`gsplat/gsplat/utils.py:248,301,302`) -- and returns raw results.
It is the GPU-side pin for everything the reference's own tests leave unpinned:
2-D projection, rasterize-sum forward/backward, projection backward, and the reference's fit
iteration speed (`bench.py` reports it as `ref_cuda`).
"""
from __future__ import annotations

import importlib.util
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_mods = {}


def available(variant: str = "o3") -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", variant, f"gsplat_ref_{variant}.so"))


def load(variant: str = "o3"):
    if variant not in _mods:
        path = os.path.join(HERE, "_ref", variant, f"gsplat_ref_{variant}.so")
        spec = importlib.util.spec_from_file_location(f"gsplat_ref_{variant}", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mods[variant] = mod
    return _mods[variant]


def tile_bounds(H, W):
    return ((W + 15) // 16, (H + 15) // 16, 1)


def project_cov(C, means2d, cov2d, H, W, clip_coe=3.0, radius_clip=1.0):
    # project_gaussians_2d_covariance.py:85-102
    return C.project_gaussians_2d_covariance_forward(means2d.shape[0], clip_coe, means2d, cov2d, H, W,
                                                     tile_bounds(H, W), 0.01, radius_clip, False)


def project_chol(C, means2d, L, H, W, radius_clip=1.0):
    # project_gaussians_2d.py:86-97
    return C.project_gaussians_2d_forward(means2d.shape[0], 3.0, means2d, L, H, W, tile_bounds(H, W), 0.01,
                                          radius_clip, False)


def project_rs(C, means2d, scales, rot, H, W, radius_clip=1.0):
    # project_gaussians_2d_scale_rot.py:92-104
    return C.project_gaussians_2d_scale_rot_forward(means2d.shape[0], 3.0, means2d, scales, rot, H, W,
                                                    tile_bounds(H, W), 0.01, radius_clip, False)


def bin_and_sort(C, xys, depths, radii, num_tiles_hit, H, W, radius_clip=1.0):
    # utils.py:231-311
    tb = tile_bounds(H, W)
    cum = torch.cumsum(num_tiles_hit, dim=0, dtype=torch.int32)
    I = int(cum[-1].item())
    ids, gids = C.map_gaussian_to_intersects(xys.shape[0], I, xys, depths, radii, cum, tb, radius_clip, False)
    ids_s, perm = torch.sort(ids)
    gids_s = torch.gather(gids, 0, perm)
    bins = C.get_tile_bin_edges(I, ids_s)
    return I, cum, ids, gids, ids_s, gids_s, bins


def rasterize_fwd(C, gids_s, bins, xys, conics, colors, opacity, H, W):
    # rasterize_sum_plus.py:139-151
    bg = torch.ones(3, device=xys.device)
    return C.rasterize_sum_plus_forward(tile_bounds(H, W), (16, 16, 1), (W, H, 1), gids_s, bins, xys, conics,
                                        colors, opacity, bg, False)


def rasterize_bwd(C, gids_s, bins, xys, conics, colors, opacity, final_Ts, final_idx, v_out, H, W):
    # rasterize_sum_plus.py:209-225
    bg = torch.ones(3, device=xys.device)
    return C.rasterize_sum_plus_backward(H, W, 16, 16, gids_s, bins, xys, conics, colors, opacity, bg, final_Ts,
                                         final_idx, v_out.contiguous(), torch.zeros_like(v_out[..., 0]))


class RefTrainer:
    """The reference's train_iter (models/gaussianimage_covariance.py:187-259) on its own extension:
    same op sequence, torch autograd glue, torch.optim.Adam(eps=1e-15) + StepLR(20000, 0.5), the
    cumsum `.item()` sync and the per-iteration PSNR `.item()` sync included."""

    def __init__(self, variant, xyz, cov, bound, rgb, gt_chw, lr=0.018):
        self.C = load(variant)
        dev = gt_chw.device
        self.xyz = torch.nn.Parameter(xyz.clone().to(dev))
        self.cov = torch.nn.Parameter(cov.clone().to(dev))
        self.rgb = torch.nn.Parameter(rgb.clone().to(dev))
        self.bound = bound.clone().to(dev)
        self.opacity = torch.ones(xyz.shape[0], 1, device=dev)
        self.gt = gt_chw
        self.H, self.W = gt_chw.shape[-2:]
        groups = [{"params": [self.xyz], "lr": lr}, {"params": [self.rgb], "lr": lr}, {"params": [self.cov], "lr": lr}]
        self.opt = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        self.sched = torch.optim.lr_scheduler.StepLR(self.opt, step_size=20000, gamma=0.5)
        C, H, W = self.C, self.H, self.W

        class Project(torch.autograd.Function):
            @staticmethod
            def forward(ctx, means, cov2d):
                out = project_cov(C, means, cov2d, H, W)
                ctx.save_for_backward(means, cov2d, out[2], out[3])
                return out

            @staticmethod
            def backward(ctx, v_xys, v_depths, v_radii, v_conics, v_nth):
                means, cov2d, radii, conics = ctx.saved_tensors
                _, v_mean, v_L = C.project_gaussians_2d_covariance_backward(
                    means.shape[0], means, cov2d, H, W, radii, conics, v_xys.contiguous(), v_depths,
                    v_conics.contiguous())
                return v_mean, v_L

        class Raster(torch.autograd.Function):
            @staticmethod
            def forward(ctx, xys, depths, radii, conics, nth, colors, opacity):
                I, cum, ids, gids, ids_s, gids_s, bins = bin_and_sort(C, xys, depths, radii, nth, H, W)
                img, Ts, fidx = rasterize_fwd(C, gids_s, bins, xys, conics, colors, opacity, H, W)
                ctx.save_for_backward(gids_s, bins, xys, conics, colors, opacity, Ts, fidx)
                return img

            @staticmethod
            def backward(ctx, v_img):
                gids_s, bins, xys, conics, colors, opacity, Ts, fidx = ctx.saved_tensors
                v_xy, v_conic, v_col, v_op = rasterize_bwd(C, gids_s, bins, xys, conics, colors, opacity, Ts, fidx,
                                                           v_img, H, W)
                return v_xy, None, None, v_conic, None, v_col, v_op

        self._project, self._raster = Project, Raster

    def forward(self):
        xys, depths, radii, conics, nth = self._project.apply(self.xyz, self.cov + self.bound)
        img = self._raster.apply(xys, depths, radii, conics, nth, self.rgb, self.opacity)
        img = torch.clamp(img, 0, 1)
        return img.view(-1, self.H, self.W, 3).permute(0, 3, 1, 2).contiguous()

    def train_iter(self):
        import math

        image = self.forward()
        loss = torch.nn.functional.mse_loss(image, self.gt)
        loss.backward()
        with torch.no_grad():
            mse = torch.nn.functional.mse_loss(image, self.gt)
            psnr = 10 * math.log10(1.0 / mse.item())
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        self.sched.step()
        return loss, psnr
