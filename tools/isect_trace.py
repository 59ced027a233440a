#!/usr/bin/env python
"""`python tools/isect_trace.py [workload]`: how the scene of a fit grows -- num_intersects, algorithmic
pairs, per-tile list length (mean / p99 / max) and the back-to-back step time at a few iterations.
Chooses the pre-roll iteration bench.py measures at.  Prints JSON lines (not bench values)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gaussianimage_plus_b200 import synth
from gaussianimage_plus_b200.fit import GaussianImageFitter

name = sys.argv[1] if len(sys.argv) > 1 else "kodak_5000"
H, W, N = synth.CONFIGS[name]
xyz, cov, bound, rgb = synth.init_covariance_model(N, H, W, seed=3047, colors="zeros")
gt_u8 = np.round(synth.target_image(H, W) * 255.0).astype(np.uint8)
fit = GaussianImageFitter(N, H, W)
for dst, src in ((fit._xyz, xyz), (fit._cov2d, cov), (fit.cholesky_bound, bound), (fit._features_dc, rgb)):
    dst.copy_(torch.from_numpy(src))
fit.set_target(torch.from_numpy(gt_u8))
done = 0
for target in (30, 130, 300, 1000, 2000, 3000, 5000, 10000):
    fit.train_iters(target - done)
    done = target
    torch.cuda.synchronize()
    fit.ensure_capacity()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fit.train_iters(200)
    e1.record()
    torch.cuda.synchronize()
    done += 200
    st = fit.stats()
    cnt = (fit.tile_bins[:, 1] - fit.tile_bins[:, 0]).clamp(min=0).float()
    print(json.dumps({"iteration": done, "num_intersects": st["num_intersects"], "psnr": round(st["psnr"], 3),
                      "per_tile_mean": round(float(cnt.mean()), 2), "per_tile_p99": float(cnt.quantile(0.99)),
                      "per_tile_max": float(cnt.max()), "pairs": float(cnt.clamp(max=256).sum()) * 256,
                      "us_per_step_warm": round(e0.elapsed_time(e1) / 200 * 1e3, 2)}), flush=True)
