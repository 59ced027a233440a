#!/usr/bin/env python
"""Runs every kernel path of the fit step on a -DGI2D_DEBUG_CHECKS build (tools/build_variant.sh dbg
"-DGI2D_DEBUG_CHECKS"; GI2D_LIB=gaussianimage_plus_b200/csrc/build/libgi2d_dbg.so) and reports the number of
violated device-side invariants (stats[15]): tile indices inside the band, placement slots inside the tile's
range, in-tile ranks a permutation, keys in the right tile, cursor == count when the rasterizer takes over.
(compute-sanitizer is closed on the development pool.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gaussianimage_plus_b200 import synth
from gaussianimage_plus_b200.fit import GaussianImageFitter

total = 0.0


def run(N, H, W, steps, loss="L2", scale=1.0, graph=False, tile_rows=None):
    global total
    xyz, cov, bound, rgb = synth.init_covariance_model(N, H, W, seed=1, colors="rand", cov_scale=scale)
    gt = torch.from_numpy(synth.target_image(H, W, seed=1))
    fit = GaussianImageFitter(N, H, W, use_graph=graph, loss_type=loss, tile_rows=tile_rows)
    for dst, src in ((fit._xyz, xyz), (fit._cov2d, cov), (fit.cholesky_bound, bound), (fit._features_dc, rgb)):
        dst.copy_(torch.from_numpy(src))
    fit.set_target(gt)
    for i in range(steps):
        fit.train_iter(want_error_map=(i == 1))
    fit.forward()
    torch.cuda.synchronize()
    bad = float(fit.stats_buf[15].item())
    total += bad
    print(f"N={N} {W}x{H} loss={loss} scale={scale} graph={graph} rows={tile_rows}: violations={bad:.0f} "
          f"I={fit.stats()['num_intersects']} psnr={fit.stats()['psnr']:.3f}")


run(400, 96, 128, 30)
run(5000, 512, 768, 200, graph=True)
run(400, 100, 130, 20, loss="Fusion2")            # ragged edges, SSIM split
run(2000, 800, 800, 20)                           # 2500 tiles: device-wide scan path
run(3000, 64, 64, 10, scale=3.0)                  # > 256 entries per tile: full in-tile sort
run(20000, 1356, 2040, 50, graph=True)
run(5000, 512, 768, 20, tile_rows=(8, 20))        # a band (multi-GPU split)
print("TOTAL VIOLATIONS", total)
sys.exit(1 if total else 0)
