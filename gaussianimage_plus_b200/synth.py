"""Deterministic synthetic workloads of the shapes BASELINE.json names.

numpy only (no device work): shared by bench.py and the tests so that the GPU path, the CPU
oracle and the reference's CUDA extension all see byte-identical inputs.
"""
from __future__ import annotations

import math

import numpy as np

CONFIGS = {
    # name: (H, W, N)
    "kodak_2500": (512, 768, 2500),      # configs[0]
    "kodak_5000": (512, 768, 5000),      # configs[1]: the headline (768x512 @ 5k Gaussians)
    "div2k_20000": (1356, 2040, 20000),  # configs[2]
    "big_1m": (8192, 8192, 1000000),     # configs[4]
}


def target_image(H: int, W: int, seed: int = 3047) -> np.ndarray:
    """f32[H,W,3] in [0,1]: 64 anisotropic blobs + 8 half-plane steps + 2% noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.zeros((H, W, 3), np.float32)
    for _ in range(64):
        cx, cy = rng.uniform(0, W), rng.uniform(0, H)
        sx, sy = rng.uniform(0.02, 0.15) * W, rng.uniform(0.02, 0.15) * H
        th = rng.uniform(0, math.pi)
        c, s = math.cos(th), math.sin(th)
        u = (xx - cx) * c + (yy - cy) * s
        v = -(xx - cx) * s + (yy - cy) * c
        w = np.exp(-0.5 * ((u / sx) ** 2 + (v / sy) ** 2)).astype(np.float32)
        img += w[..., None] * rng.uniform(0.1, 0.6, 3).astype(np.float32)
    for _ in range(8):
        nx, ny = rng.normal(size=2)
        d = rng.uniform(-0.3, 0.3) * max(H, W)
        mask = ((xx - W / 2) * nx + (yy - H / 2) * ny > d).astype(np.float32)
        img += mask[..., None] * rng.uniform(-0.25, 0.25, 3).astype(np.float32)
    img += rng.uniform(-0.02, 0.02, img.shape).astype(np.float32)
    return np.clip(img, 0.0, 1.0).astype(np.float32)


def init_covariance_model(N: int, H: int, W: int, seed: int = 3047, colors: str = "zeros", cov_scale: float = 1.0):
    """Reference initialisation of GaussianImage_Covariance (models/gaussianimage_covariance.py:52-66):
    xyz = U(0,1)*(W,H), _cov2d = U(0,1)^3, cholesky_bound = (lp,0,lp) with lp = min(HW/(9 pi N), 300),
    features_dc = 0 (or U(0,1) for kernel micro-benchmarks).  `cov_scale` > 1 emulates the wider
    Gaussians of a mid-training state."""
    rng = np.random.default_rng(seed)
    xyz = (rng.random((N, 2), np.float32) * np.array([W, H], np.float32)).astype(np.float32)
    cov = rng.random((N, 3), np.float32)
    lp = min(H * W / (9 * math.pi * N), 300) * cov_scale
    bound = np.tile(np.array([lp, 0, lp], np.float32), (N, 1))
    rgb = np.zeros((N, 3), np.float32) if colors == "zeros" else rng.random((N, 3), np.float32)
    return xyz, cov, bound, rgb


def cholesky_inputs(N: int, H: int, W: int, seed: int = 3047):
    """configs[0]: Cholesky parameterisation; means in [-1,1] (tanh of the raw parameter,
    models/gaussianimage_cholesky.py:137), L = U(0,1)^3 + (0.5, 0, 0.5)."""
    rng = np.random.default_rng(seed)
    means = np.tanh(np.arctanh(np.clip(2 * rng.random((N, 2)) - 1, -0.999, 0.999))).astype(np.float32)
    L = (rng.random((N, 3)) + np.array([0.5, 0, 0.5])).astype(np.float32) * 3.0
    colors = rng.random((N, 3), np.float32)
    return means, L.astype(np.float32), colors


def scale_rot_inputs(N: int, H: int, W: int, seed: int = 3047):
    rng = np.random.default_rng(seed)
    means = (rng.random((N, 2), np.float32) * np.array([W, H], np.float32)).astype(np.float32)
    scales = (rng.random((N, 2)) * 6 + 0.5).astype(np.float32)
    rot = (rng.random((N, 1)) * 2 * math.pi).astype(np.float32)
    colors = rng.random((N, 3), np.float32)
    return means, scales, rot, colors


def target_image_u8_torch(H: int, W: int, seed: int = 3047, device="cuda"):
    """u8[H,W,3] on `device`: the same family of targets as target_image() (64 anisotropic blobs + 8 half-plane
    steps + 2 % noise, clamp, 8-bit), evaluated with torch on the device -- for the 8192^2 workload, whose 67 M
    pixels take minutes in numpy on a host core.  The blob / step parameters come from the same numpy stream as
    target_image(); only the per-pixel noise differs (torch generator).  Deterministic for a given seed."""
    import torch

    rng = np.random.default_rng(seed)
    dev = torch.device(device)
    yy = torch.arange(H, device=dev, dtype=torch.float32).view(H, 1)
    xx = torch.arange(W, device=dev, dtype=torch.float32).view(1, W)
    img = torch.zeros(H, W, 3, device=dev)
    for _ in range(64):
        cx, cy = rng.uniform(0, W), rng.uniform(0, H)
        sx, sy = rng.uniform(0.02, 0.15) * W, rng.uniform(0.02, 0.15) * H
        th = rng.uniform(0, math.pi)
        c, s = math.cos(th), math.sin(th)
        col = torch.tensor(rng.uniform(0.1, 0.6, 3).astype(np.float32), device=dev)
        u = (xx - cx) * c + (yy - cy) * s
        v = -(xx - cx) * s + (yy - cy) * c
        img += torch.exp(-0.5 * ((u / sx) ** 2 + (v / sy) ** 2)).unsqueeze(-1) * col
    for _ in range(8):
        nx, ny = rng.normal(size=2)
        d = rng.uniform(-0.3, 0.3) * max(H, W)
        col = torch.tensor(rng.uniform(-0.25, 0.25, 3).astype(np.float32), device=dev)
        img += (((xx - W / 2) * nx + (yy - H / 2) * ny) > d).float().unsqueeze(-1) * col
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    img += (torch.rand(H, W, 3, device=dev, generator=g) - 0.5) * 0.04
    return (img.clamp_(0.0, 1.0) * 255.0).round_().to(torch.uint8)
