#!/usr/bin/env python
"""tools/prof_dropin.py: cProfile of the reference's train_iter protocol on the drop-in operators (host side).
Not a bench value."""
import cProfile
import math
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gaussianimage_plus_b200 as pkg
from gaussianimage_plus_b200 import synth

pkg.install_as_gsplat()
from gsplat.project_gaussians_2d_covariance import project_gaussians_2d_covariance
from gsplat.rasterize_sum_plus import rasterize_gaussians_plus

dev = "cuda:0"
H, W, N = synth.CONFIGS["kodak_5000"]
xyz, cov, bound, rgb = synth.init_covariance_model(N, H, W, seed=3047, colors="zeros")
gt = synth.target_image(H, W)
gt_chw = torch.from_numpy(gt).to(dev).permute(2, 0, 1).unsqueeze(0).contiguous()
p_xyz, p_cov, p_rgb = (torch.nn.Parameter(torch.from_numpy(a).to(dev)) for a in (xyz, cov, rgb))
bnd = torch.from_numpy(bound).to(dev)
opacity = torch.ones(N, 1, device=dev)
opt = torch.optim.Adam([{"params": [p_xyz], "lr": 0.018}, {"params": [p_rgb], "lr": 0.018},
                        {"params": [p_cov], "lr": 0.018}], lr=0.0, eps=1e-15)
tb = ((W + 15) // 16, (H + 15) // 16, 1)


def train_iter():
    xys, depths, radii, conics, nth = project_gaussians_2d_covariance(p_xyz, p_cov + bnd, H, W, tb)
    out = rasterize_gaussians_plus(xys, depths, radii, conics, nth, p_rgb, opacity, H, W, 16, 16)
    image = torch.clamp(out, 0, 1).view(-1, H, W, 3).permute(0, 3, 1, 2).contiguous()
    loss = torch.nn.functional.mse_loss(image, gt_chw)
    loss.backward()
    with torch.no_grad():
        psnr = 10 * math.log10(1.0 / torch.nn.functional.mse_loss(image, gt_chw).item())
    opt.step()
    opt.zero_grad(set_to_none=True)
    return psnr


for _ in range(50):
    train_iter()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    train_iter()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
