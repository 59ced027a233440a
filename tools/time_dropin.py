#!/usr/bin/env python
"""it/s of the reference's train_iter protocol (models/gaussianimage_covariance.py:187-259: torch autograd,
torch.optim.Adam, per-iteration PSNR .item()) when ONLY the operators are swapped -- the drop-in `gsplat`
package of this repo -- next to the unmodified reference extension on the same GPU."""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gaussianimage_plus_b200 as pkg
from gaussianimage_plus_b200 import synth

pkg.install_as_gsplat()
from gsplat.project_gaussians_2d_covariance import project_gaussians_2d_covariance  # noqa: E402
from gsplat.rasterize_sum_plus import rasterize_gaussians_plus  # noqa: E402

H, W, N = synth.CONFIGS["kodak_5000"]
dev = torch.device("cuda:0")
xyz, cov, bound, rgb = synth.init_covariance_model(N, H, W, seed=3047, colors="zeros")
gt = torch.from_numpy(synth.target_image(H, W)).to(dev).permute(2, 0, 1).unsqueeze(0).contiguous()
p_xyz, p_cov, p_rgb = (torch.nn.Parameter(torch.from_numpy(a).to(dev)) for a in (xyz, cov, rgb))
bnd = torch.from_numpy(bound).to(dev)
opacity = torch.ones(N, 1, device=dev)
opt = torch.optim.Adam([{"params": [p_xyz], "lr": 0.018}, {"params": [p_rgb], "lr": 0.018},
                        {"params": [p_cov], "lr": 0.018}], lr=0.0, eps=1e-15)
sched = torch.optim.lr_scheduler.StepLR(opt, step_size=20000, gamma=0.5)
tb = ((W + 15) // 16, (H + 15) // 16, 1)


def train_iter():
    xys, depths, radii, conics, nth = project_gaussians_2d_covariance(p_xyz, p_cov + bnd, H, W, tb)
    out = rasterize_gaussians_plus(xys, depths, radii, conics, nth, p_rgb, opacity, H, W, 16, 16)
    image = torch.clamp(out, 0, 1).view(-1, H, W, 3).permute(0, 3, 1, 2).contiguous()
    loss = torch.nn.functional.mse_loss(image, gt)
    loss.backward()
    with torch.no_grad():
        psnr = 10 * math.log10(1.0 / torch.nn.functional.mse_loss(image, gt).item())
    opt.step()
    opt.zero_grad(set_to_none=True)
    sched.step()
    return psnr


for _ in range(30):
    train_iter()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    psnr = train_iter()
torch.cuda.synchronize()
ours = 300 / (time.perf_counter() - t0)
res = {"dropin_operators_it_s": ours, "psnr_after_330": psnr}
try:
    from oracle import ref_cuda

    if ref_cuda.available("fastmath"):
        tr = ref_cuda.RefTrainer("fastmath", *(torch.from_numpy(a) for a in (xyz, cov, bound, rgb)), gt)
        for _ in range(30):
            tr.train_iter()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(300):
            _, rp = tr.train_iter()
        torch.cuda.synchronize()
        res.update({"reference_extension_it_s": 300 / (time.perf_counter() - t0), "reference_psnr_after_330": rp})
except Exception as e:  # noqa: BLE001
    res["reference_extension"] = repr(e)[:200]
print(json.dumps(res))
