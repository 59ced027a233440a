#!/usr/bin/env python
"""Small driver for compute-sanitizer: touches every kernel path of the fit step once (smem-scan and
device-scan binning, the > 256-per-tile sort path, SSIM split, render, flush, prune, error map)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gaussianimage_plus_b200 import synth
from gaussianimage_plus_b200.fit import GaussianImageFitter


def run(N, H, W, steps, loss="L2", scale=1.0, u8=False):
    xyz, cov, bound, rgb = synth.init_covariance_model(N, H, W, seed=1, colors="rand", cov_scale=scale)
    gt = torch.from_numpy(synth.target_image(H, W, seed=1))
    if u8:
        gt = (gt * 255).round().to(torch.uint8)
    fit = GaussianImageFitter(N, H, W, use_graph=False, loss_type=loss)
    for dst, src in ((fit._xyz, xyz), (fit._cov2d, cov), (fit.cholesky_bound, bound), (fit._features_dc, rgb)):
        dst.copy_(torch.from_numpy(src))
    fit.set_target(gt)
    for i in range(steps):
        fit.train_iter(want_error_map=(i == 1))
    fit.forward()
    fit.non_semi_definite_prune()
    torch.cuda.synchronize()
    print(N, H, W, loss, fit.stats()["psnr"], fit.stats()["num_intersects"])


run(400, 96, 128, 3)
run(400, 100, 130, 3, loss="Fusion2", u8=True)     # ragged edges, SSIM split, 8-bit target
run(2000, 800, 800, 2)                             # 2500 tiles: device-wide scan path
run(3000, 64, 64, 2, scale=3.0)                    # > 256 entries per tile: full in-tile sort
