#!/usr/bin/env python
"""Whole-fit comparison at BASELINE.json's headline size (768x512, 5000 Gaussians, L2, Adam + StepLR, fixed
Gaussian count): GaussianImageFitter.fit vs the UNMODIFIED reference CUDA extension driven with the reference's
train_iter protocol (oracle/ref_cuda.RefTrainer) -- PSNR at equal iterations and wall-clock time to get there.
Prints one JSON line.  `python tools/full_fit_compare.py [iterations]`"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gaussianimage_plus_b200 import synth
from gaussianimage_plus_b200.fit import GaussianImageFitter
from oracle import ref_cuda

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
marks = [m for m in (100, 300, 1000, 2000, 5000, 10000, 20000, 50000) if m <= iters]
H, W, N = synth.CONFIGS["kodak_5000"]
dev = torch.device("cuda:0")
xyz, cov, bound, rgb = synth.init_covariance_model(N, H, W, seed=3047, colors="zeros")
gt_u8 = np.round(synth.target_image(H, W) * 255.0).astype(np.uint8)
gt = gt_u8.astype(np.float32) / np.float32(255.0)

# ---- ours
fit = GaussianImageFitter(N, H, W, device=dev)
for dst, src in ((fit._xyz, xyz), (fit._cov2d, cov), (fit.cholesky_bound, bound), (fit._features_dc, rgb)):
    dst.copy_(torch.from_numpy(src))
fit.set_target(torch.from_numpy(gt_u8))
fit.train_iters(8)                       # loads the kernels, captures the graphs (not part of either clock)
fit.reset_stats(0)
for dst, src in ((fit._xyz, xyz), (fit._cov2d, cov), (fit._features_dc, rgb)):
    dst.copy_(torch.from_numpy(src))
for d in (fit._t_m, fit._t_v):
    for t in d.values():
        t.zero_()
ours, done = {}, 0
torch.cuda.synchronize()
t0 = time.perf_counter()
for m in marks:
    fit.train_iters(m - done)
    done = m
    st = fit.stats()                      # synchronises: PSNR of iteration m
    ours[m] = {"psnr": st["psnr"], "seconds": time.perf_counter() - t0}
fit.sync_params()
torch.cuda.synchronize()
ours_total = time.perf_counter() - t0
best = fit.stats()

# ---- the reference extension, its own protocol (per-iteration .item() syncs included)
ref = {}
if ref_cuda.available("fastmath"):
    gt_chw = torch.from_numpy(gt).to(dev).permute(2, 0, 1).unsqueeze(0).contiguous()
    tr = ref_cuda.RefTrainer("fastmath", *(torch.from_numpy(a) for a in (xyz, cov, bound, rgb)), gt_chw)
    warm = ref_cuda.RefTrainer("fastmath", *(torch.from_numpy(a) for a in (xyz, cov, bound, rgb)), gt_chw)
    for _ in range(5):
        warm.train_iter()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(1, iters + 1):
        _, psnr = tr.train_iter()
        if it in marks:
            ref[it] = {"psnr": psnr, "seconds": time.perf_counter() - t0}
    torch.cuda.synchronize()
    ref_total = time.perf_counter() - t0
else:
    ref_total = None
print(json.dumps({"workload": f"kodak_5000: {W}x{H}, {N} Gaussians, L2, fixed count, {iters} iterations",
                  "ours": ours, "ours_total_s": ours_total, "ours_best_psnr": best["best_psnr"],
                  "reference_cuda_fastmath": ref, "reference_total_s": ref_total,
                  "speedup_wall_clock": (ref_total / ours_total) if ref_total else None}))
