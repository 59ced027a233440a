"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU oracle of the image losses of the reference, `loss_fn` in /root/reference/models/utils.py:60-80,
in float64 torch with autograd.  `ssim` is a THIRD-PARTY dependency that is not vendored in
/root/reference and is unpinned there (requirements.txt: "pytorch-msssim"; not installed in this
image, no network): what follows restates the published algorithm of pytorch_msssim 1.0.0
(`pytorch_msssim/ssim.py`: `_fspecial_gauss_1d`, `gaussian_filter`, `_ssim`, `ssim`), anchored on the
reference's call sites `ssim(pred, target, data_range=1, size_average=True)` (models/utils.py:69-73).
Parity for this function is therefore "unpinned" (no golden vector in the reference, the package is
absent); its own sanity anchors are tested in tests/test_oracle_golden.py (ssim(x,x) == 1, symmetry,
the window sums to 1, a hand-computed constant-image value).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def fspecial_gauss_1d(size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    """pytorch_msssim `_fspecial_gauss_1d`: float32 arithmetic, then normalised."""
    coords = torch.arange(size, dtype=torch.float)
    coords -= size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    g /= g.sum()
    return g


def gaussian_filter(x: torch.Tensor, win: torch.Tensor) -> torch.Tensor:
    """pytorch_msssim `gaussian_filter`: separable, groups = channels, NO padding (valid)."""
    C = x.shape[1]
    k = win.to(x).view(1, 1, -1).repeat(C, 1, 1)
    out = F.conv2d(x, k.unsqueeze(-1), groups=C)    # along H
    out = F.conv2d(out, k.unsqueeze(-2), groups=C)  # along W
    return out


def ssim(X: torch.Tensor, Y: torch.Tensor, data_range: float = 1.0, K=(0.01, 0.03)) -> torch.Tensor:
    """pytorch_msssim `ssim(X, Y, data_range, size_average=True)` for [B,C,H,W] inputs."""
    win = fspecial_gauss_1d()
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = gaussian_filter(X, win), gaussian_filter(Y, win)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    sigma1_sq = gaussian_filter(X * X, win) - mu1_sq
    sigma2_sq = gaussian_filter(Y * Y, win) - mu2_sq
    sigma12 = gaussian_filter(X * Y, win) - mu1_mu2
    cs_map = (2 * sigma12 + C2) / (sigma1_sq + sigma2_sq + C2)
    ssim_map = ((2 * mu1_mu2 + C1) / (mu1_sq + mu2_sq + C1)) * cs_map
    return torch.flatten(ssim_map, 2).mean(-1).mean()


def _ssim_and_cs(X, Y, data_range=1.0, K=(0.01, 0.03), win_size: int = 11):
    """pytorch_msssim `_ssim(..., size_average=False)`: per-channel means of the SSIM map and of the cs map."""
    win = fspecial_gauss_1d(win_size)
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = gaussian_filter(X, win), gaussian_filter(Y, win)
    sigma1_sq = gaussian_filter(X * X, win) - mu1 * mu1
    sigma2_sq = gaussian_filter(Y * Y, win) - mu2 * mu2
    sigma12 = gaussian_filter(X * Y, win) - mu1 * mu2
    cs_map = (2 * sigma12 + C2) / (sigma1_sq + sigma2_sq + C2)
    ssim_map = ((2 * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs_map
    return torch.flatten(ssim_map, 2).mean(-1), torch.flatten(cs_map, 2).mean(-1)


MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def ms_ssim(X: torch.Tensor, Y: torch.Tensor, data_range: float = 1.0, win_size: int = 11) -> torch.Tensor:
    """pytorch_msssim `ms_ssim(X, Y, data_range, size_average=True, win_size=...)` (the evaluation metric of
    train.py:190 and the `1 - ms_ssim` term of Fusion4 / Fusion_hinerv, models/utils.py:76-79): 5 levels,
    avg_pool2d(kernel 2, padding = size % 2) between them, relu on cs / ssim, weighted product; the window is
    `_fspecial_gauss_1d(win_size, 1.5)` (win_sigma keeps its default when only win_size is given)."""
    assert min(X.shape[-2:]) > (win_size - 1) * 2 ** 4
    w = X.new_tensor(MS_WEIGHTS)
    mcs = []
    for i in range(5):
        ssim_c, cs = _ssim_and_cs(X, Y, data_range, win_size=win_size)
        if i < 4:
            mcs.append(torch.relu(cs))
            pad = [s % 2 for s in X.shape[2:]]
            X = F.avg_pool2d(X, kernel_size=2, padding=pad)
            Y = F.avg_pool2d(Y, kernel_size=2, padding=pad)
    stack = torch.stack(mcs + [torch.relu(ssim_c)], dim=0)
    return torch.prod(stack ** w.view(-1, 1, 1), dim=0).mean()


def loss_fn(pred: torch.Tensor, target: torch.Tensor, loss_type: str = "L2", lambda_value: float = 0.7):
    """models/utils.py:60-80."""
    if loss_type == "Fusion4":
        return lambda_value * F.l1_loss(pred, target) + (1 - lambda_value) * (1 - ms_ssim(pred, target))
    if loss_type == "Fusion_hinerv":
        return lambda_value * F.l1_loss(pred, target) + (1 - lambda_value) * (1 - ms_ssim(pred, target, win_size=5))
    if loss_type == "L2":
        return F.mse_loss(pred, target)
    if loss_type == "L1":
        return F.l1_loss(pred, target)
    if loss_type == "SSIM":
        return 1 - ssim(pred, target)
    if loss_type == "Fusion1":
        return lambda_value * F.mse_loss(pred, target) + (1 - lambda_value) * (1 - ssim(pred, target))
    if loss_type == "Fusion2":
        return lambda_value * F.l1_loss(pred, target) + (1 - lambda_value) * (1 - ssim(pred, target))
    if loss_type == "Fusion3":
        return lambda_value * F.mse_loss(pred, target) + (1 - lambda_value) * F.l1_loss(pred, target)
    raise ValueError(loss_type)


def loss_and_grad(render_hwc, gt_hwc, loss_type: str, lambda_value: float = 0.7, dtype=torch.float64):
    """What autograd computes between the rasterizer output and the loss in
    gaussianimage_covariance.py:210,252-253: clamp(0,1) -> [1,3,H,W] -> loss_fn -> backward.
    Returns (loss, d loss / d render [H,W,3], mean ssim)."""
    r = torch.as_tensor(render_hwc).to(dtype).clone().requires_grad_(True)
    g = torch.as_tensor(gt_hwc).to(dtype)
    pred = torch.clamp(r, 0, 1).permute(2, 0, 1).unsqueeze(0)
    tgt = g.permute(2, 0, 1).unsqueeze(0)
    loss = loss_fn(pred, tgt, loss_type, lambda_value)
    loss.backward()
    with torch.no_grad():
        s = float(ssim(pred, tgt))
    return float(loss.detach()), r.grad.detach(), s
