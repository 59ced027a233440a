"""Generates tests/golden/*.npz by IMPORTING the reference's own pure-PyTorch helpers
(`/root/reference/gsplat/gsplat/_torch_impl.py`) on the exact fixtures of the reference's tests:

  gsplat/tests/test_map_gaussians.py:9-73       isect_ids / gaussian_ids   (exact)
  gsplat/tests/test_get_tile_bin_edges.py:9-81  tile_bins                  (exact)
  gsplat/tests/test_cov2d_bounds.py:9-35        conic / radius             (assert_close)

Those three are the ONLY known-answer tests the reference holds for this path;
the fixtures are seed 42, 100 points, 512x512, inputs produced by the 3-D
`_torch_impl.project_gaussians_forward`.  The reference cannot travel to the GPU box, so the
vectors are committed and this script documents how they were made.

Run here (container with /root/reference):   python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def load_torch_impl():
    spec = importlib.util.spec_from_file_location(
        "ref_torch_impl", "/root/reference/gsplat/gsplat/_torch_impl.py"
    )
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def binning_fixture(ti):
    # verbatim protocol of gsplat/tests/test_map_gaussians.py:13-65 on CPU
    torch.manual_seed(42)
    num_points = 100
    means3d = torch.randn((num_points, 3))
    scales = torch.randn((num_points, 3))
    glob_scale = 0.3
    quats = torch.randn((num_points, 4))
    quats /= torch.linalg.norm(quats, dim=-1, keepdim=True)
    viewmat = torch.eye(4)
    projmat = torch.eye(4)
    fx, fy = 3.0, 3.0
    H, W = 512, 512
    clip_thresh = 0.01
    tile_bounds = ((W + 15) // 16, (H + 15) // 16, 1)
    (_cov3d, xys, depths, radii, conics, num_tiles_hit, masks) = ti.project_gaussians_forward(
        means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, (H, W), tile_bounds, clip_thresh
    )
    xys, depths, radii, conics, num_tiles_hit = (t[masks] for t in (xys, depths, radii, conics, num_tiles_hit))
    num_points = num_points - torch.count_nonzero(~masks).item()
    cum = torch.cumsum(num_tiles_hit, dim=0, dtype=torch.int32)
    num_intersects = cum[-1].item()
    depths = depths.contiguous()
    isect_ids, gaussian_ids = ti.map_gaussian_to_intersects(num_points, xys, depths, radii, cum, tile_bounds)
    isect_sorted, perm = torch.sort(isect_ids)
    gids_sorted = torch.gather(gaussian_ids, 0, perm)
    tile_bins = ti.get_tile_bin_edges(num_intersects, isect_sorted)
    return dict(
        H=H, W=W, tile_bounds=np.array(tile_bounds, np.int32), num_points=num_points,
        xys=xys.numpy().astype(np.float32), depths=depths.numpy().astype(np.float32),
        radii=radii.numpy().astype(np.int32), num_tiles_hit=num_tiles_hit.numpy().astype(np.int32),
        cum_tiles_hit=cum.numpy(), num_intersects=num_intersects,
        isect_ids=isect_ids.numpy(), gaussian_ids=gaussian_ids.numpy(),
        isect_ids_sorted=isect_sorted.numpy(), gaussian_ids_sorted=gids_sorted.numpy(),
        tile_bins=tile_bins.numpy(),
    )


def cov2d_fixture(ti):
    # gsplat/tests/test_cov2d_bounds.py:13-31
    torch.manual_seed(42)
    n = 100
    _covs2d = torch.rand((n, 2, 2), dtype=torch.float32)
    covs2d = torch.stack(
        [torch.triu(_covs2d)[:, 0, 0], torch.triu(_covs2d)[:, 0, 1], torch.triu(_covs2d)[:, 1, 1]], dim=-1
    )
    conic, radii, mask = ti.compute_cov2d_bounds(_covs2d)
    return dict(covs2d=covs2d.numpy(), conic=conic.numpy(), radii=radii.numpy(), mask=mask.numpy())


if __name__ == "__main__":
    if not os.path.isdir("/root/reference"):
        sys.exit("reference not present: golden vectors can only be regenerated in the build container")
    ti = load_torch_impl()
    np.savez(os.path.join(HERE, "binning_seed42.npz"), **binning_fixture(ti))
    np.savez(os.path.join(HERE, "cov2d_bounds_seed42.npz"), **cov2d_fixture(ti))
    print("wrote", os.listdir(HERE))
