#!/usr/bin/env python
"""Short driver for ncu: `python tools/prof_fit.py [workload] [warm steps] [steps] [loss]` runs un-graphed
fit steps of a BASELINE.json workload (so every kernel is an ordinary launch the profiler can pick by name).
The numbers this prints are NOT bench values."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gaussianimage_plus_b200 import synth
from gaussianimage_plus_b200.fit import GaussianImageFitter

name = sys.argv[1] if len(sys.argv) > 1 else "kodak_5000"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 300
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
loss = sys.argv[4] if len(sys.argv) > 4 else "L2"
H, W, N = synth.CONFIGS[name]
xyz, cov, bound, rgb = synth.init_covariance_model(N, H, W, seed=3047, colors="zeros")
if H * W > (1 << 23):
    gt_u8 = synth.target_image_u8_torch(H, W, device="cuda:0")
else:
    gt_u8 = torch.from_numpy(np.round(synth.target_image(H, W) * 255.0).astype(np.uint8))
fit = GaussianImageFitter(N, H, W, use_graph=False, loss_type=loss)
for dst, src in ((fit._xyz, xyz), (fit._cov2d, cov), (fit.cholesky_bound, bound), (fit._features_dc, rgb)):
    dst.copy_(torch.from_numpy(src))
fit.set_target(gt_u8)
for _ in range(warm + steps):
    fit.train_iter()
torch.cuda.synchronize()
print(fit.stats())
