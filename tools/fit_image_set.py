#!/usr/bin/env python
"""BASELINE.json configs[1]+[3] end to end: a Kodak-24-shaped image set (18 landscape 768x512 + 6 portrait
512x768, synthetic) sharded over the ranks with NO collective on the data path (`parallel.shard_images`), every
image fitted like the reference's main loop (train.py:276-340): 2500 -> 5000 Gaussians with error-driven
densification, pruning every 100 iterations, best state kept; then the compression pass of train_quantize.py
(quantisation-aware steps with colour normalisation optional) and the codec analysis; one gather of the
per-image metrics at the end (train.py:327-340 averages).

    python tools/fit_image_set.py [--images 24] [--iterations 5000] [--qat 200]
    torchrun --nproc-per-node N tools/fit_image_set.py ...
Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from gaussianimage_plus_b200 import synth
from gaussianimage_plus_b200.codec import FusedQuantizedTrainer, KernelQuantizedTrainer, QuantizedGaussianImage
from gaussianimage_plus_b200.fit import GaussianImageFitter
from gaussianimage_plus_b200.parallel import gather_metrics, shard_images

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=24)
ap.add_argument("--iterations", type=int, default=5000)
ap.add_argument("--num-points", type=int, default=2500)
ap.add_argument("--max-points", type=int, default=5000)
ap.add_argument("--grow-iter", type=int, default=1000)
ap.add_argument("--qat", type=int, default=200)
ap.add_argument("--color-norm", action="store_true")
ap.add_argument("--qat-operator-path", action="store_true", help="quantisation-aware steps on the autograd operator "
                "path instead of the kernel trainer")
ap.add_argument("--qat-torch-graph", action="store_true", help="round 1's FusedQuantizedTrainer (torch quantiser modules "
                "+ torch.optim.Adam replayed from a graph) instead of KernelQuantizedTrainer")
args = ap.parse_args()

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}")
torch.cuda.set_device(dev)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(3047)                                    # train.py:225
local = []
t_all = time.perf_counter()
for i in shard_images(args.images, world, rank):
    H, W = (768, 512) if i % 4 == 3 else (512, 768)        # 6 of 24 portrait, like Kodak
    gt_u8 = torch.from_numpy(np.round(synth.target_image(H, W, seed=100 + i) * 255).astype(np.uint8)).to(dev)
    fit = GaussianImageFitter(args.num_points, H, W, device=dev, color_norm=args.color_norm)
    fit.set_target(gt_u8)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    st = fit.fit(args.iterations, max_num_points=args.max_points, prune_iter=100, grow_iter=args.grow_iter)
    torch.cuda.synchronize(dev)
    t_fit = time.perf_counter() - t0
    fit.load_best_state()
    ms = fit.ms_ssim()
    rec = {"image": i, "HxW": f"{H}x{W}", "gaussians": fit.cur_num_points, "best_psnr": st["best_psnr"],
           "ms_ssim": ms, "fit_seconds": t_fit, "it_per_s": args.iterations / t_fit}
    if args.qat:
        if args.qat_operator_path:
            q = QuantizedGaussianImage.from_fitter(fit, best=False)
            t0 = time.perf_counter()
            q_first = None
            for _ in range(args.qat):
                _, _, _, _, qpsnr = q.train_iter_quantize(gt_u8)
                q_first = qpsnr if q_first is None else q_first
        else:
            q = (FusedQuantizedTrainer if args.qat_torch_graph else KernelQuantizedTrainer).from_fitter(fit, best=False)
            q.set_target(gt_u8)
            t0 = time.perf_counter()
            q.train_iter_quantize()
            q_first = q.psnr()                      # the first quantised forward == post-training quantisation
            for _ in range(args.qat - 1):
                q.train_iter_quantize()
            qpsnr = q.psnr()
        torch.cuda.synchronize(dev)
        enc = q.compress_wo_ec()
        dec = q.decompress_wo_ec(enc)["render"]
        mse = float(((dec[0].permute(1, 2, 0) - gt_u8.float() / 255) ** 2).mean())
        rec.update({"qat_seconds": time.perf_counter() - t0, "ptq_psnr": q_first, "qat_last_psnr": qpsnr,
                    "codec_psnr": 10 * np.log10(1 / mse),
                    "bpp": q.analysis_wo_ec(enc)["bpp"]})
    local.append((i, rec["best_psnr"], rec["fit_seconds"], rec))
allm = gather_metrics([(a, b, c) for a, b, c, _ in local])
recs = [None] * world
if world > 1:
    dist.all_gather_object(recs, [r for _, _, _, r in local])
else:
    recs = [[r for _, _, _, r in local]]
wall = time.perf_counter() - t_all
if rank == 0:
    flat = sorted((r for part in recs for r in part), key=lambda r: r["image"])
    out = {"images": args.images, "n_gpus": world, "iterations": args.iterations,
           "avg_best_psnr": float(np.mean([r["best_psnr"] for r in flat])),
           "avg_ms_ssim": float(np.mean([r["ms_ssim"] for r in flat])),
           "avg_gaussians": float(np.mean([r["gaussians"] for r in flat])),
           "avg_fit_seconds_per_image": float(np.mean([r["fit_seconds"] for r in flat])),
           "sum_fit_it_per_s": float(sum(args.iterations / r["fit_seconds"] for r in flat) / len(flat) * world),
           "wall_seconds": wall, "per_image": flat}
    if args.qat:
        out["avg_codec_psnr"] = float(np.mean([r["codec_psnr"] for r in flat]))
        out["avg_bpp"] = float(np.mean([r["bpp"] for r in flat]))
    assert [m[0] for m in allm] == list(range(args.images))
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
