#!/usr/bin/env python
"""tools/prof_fit_loop.py: cProfile of GaussianImageFitter.fit() on configs[1] as written (2500 -> 5000 Gaussians,
prune every 100, densify every 1000): where the HOST time of the loop goes.  Not a bench value."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gaussianimage_plus_b200 import synth
from gaussianimage_plus_b200.fit import GaussianImageFitter

H, W, _ = synth.CONFIGS["kodak_5000"]
xyz, cov, bound, rgb = synth.init_covariance_model(2500, H, W, seed=3047, colors="zeros")
gt_u8 = torch.from_numpy(np.round(synth.target_image(H, W) * 255.0).astype(np.uint8))
for trial in range(2):
    fit = GaussianImageFitter(2500, H, W, device="cuda:0", max_num_points=5000)
    for dst, src in ((fit._xyz, xyz), (fit._cov2d, cov), (fit.cholesky_bound, bound), (fit._features_dc, rgb)):
        dst.copy_(torch.from_numpy(src))
    fit.set_target(gt_u8)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    st = fit.fit(5000, max_num_points=5000, prune_iter=100, grow_iter=1000)
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    pr.disable()
    print("trial", trial, "enqueue+finish", t_enq, "total", time.perf_counter() - t0, "psnr", st["best_psnr"])
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
