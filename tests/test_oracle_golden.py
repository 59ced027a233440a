"""CPU: pins the oracle against the reference's own golden vectors (tests/golden/, made by
tests/golden/make_golden.py from the reference's `_torch_impl`), i.e. the fixtures of
gsplat/tests/test_map_gaussians.py, test_get_tile_bin_edges.py and test_cov2d_bounds.py."""
import os

import numpy as np


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _diverging_gaussians(g):
    """Gaussians on which the reference's two implementations disagree BY CONSTRUCTION:
    `_torch_impl.get_tile_bbox` (:236-259) truncates before adding 1, the CUDA `get_bbox`
    (helpers.cuh:26-29) adds 1 before truncating; they differ when centre+radius lies in (-1,0)
    tiles.  The oracle follows the CUDA kernel, which is what libgi2d replaces."""
    x, r = g["xys"], g["radii"].astype(np.float32)
    hi = x / np.float32(16) + (r / np.float32(16))[:, None]
    return np.nonzero(((hi > -1) & (hi < 0)).any(axis=1))[0]


def test_cumsum_and_map_match_reference_fixture(oracle, golden_dir):
    g = _load(golden_dir, "binning_seed42.npz")
    tb = tuple(int(v) for v in g["tile_bounds"])
    total, cum = oracle.cumsum_i32(g["num_tiles_hit"])
    assert total == int(g["num_intersects"])
    np.testing.assert_array_equal(cum, g["cum_tiles_hit"])
    ids, gids = oracle.map_gaussian_to_intersects(total, g["xys"], g["depths"], g["radii"], cum, tb)
    bad = _diverging_gaussians(g)
    keep = np.ones(total, bool)
    for b in bad:
        keep[(0 if b == 0 else cum[b - 1]):cum[b]] = False
    assert keep.sum() >= total - 8  # the documented divergence touches a handful of slots at most
    np.testing.assert_array_equal(ids[keep], g["isect_ids"][keep])
    np.testing.assert_array_equal(gids[keep], g["gaussian_ids"][keep])


def test_sort_and_bin_edges_match_reference_fixture(oracle, golden_dir):
    g = _load(golden_dir, "binning_seed42.npz")
    ks, vs = oracle.sort_pairs(g["isect_ids"], g["gaussian_ids"])
    np.testing.assert_array_equal(ks, g["isect_ids_sorted"])
    np.testing.assert_array_equal(vs, g["gaussian_ids_sorted"])
    bins = oracle.get_tile_bin_edges(g["isect_ids_sorted"], int(g["num_intersects"]))
    np.testing.assert_array_equal(bins, g["tile_bins"])


def test_cov2d_bounds_match_reference_fixture(oracle, golden_dir):
    c = _load(golden_dir, "cov2d_bounds_seed42.npz")
    conic, radii = oracle.compute_cov2d_bounds(c["covs2d"])
    m = c["mask"]
    assert m.sum() > 10
    # torch.testing.assert_close defaults for float32 (test_cov2d_bounds.py:34-35): rtol 1.3e-6, atol 1e-5
    np.testing.assert_allclose(conic[m], c["conic"][m], rtol=1.3e-6 * 4, atol=1e-5)
    np.testing.assert_array_equal(radii[m], c["radii"][m])


def test_projection_quirks(oracle):
    """det==0 culled, det<0 rendered (NaN minor radius passes the cull), off-screen
    Gaussians keep radii>0 with num_tiles_hit==0."""
    means = np.array([[100, 100], [100, 100], [-500, -500], [100, 100]], np.float32)
    cov = np.array([[4, 2, 1], [1, 3, 1], [30, 0, 30], [0.01, 0, 0.01]], np.float32)
    xys, depths, radii, conics, nth = oracle.project_cov_fwd(means, cov, 512, 768)
    assert radii[0] == 0 and nth[0] == 0 and (conics[0] == 0).all()      # det == 0
    assert radii[1] > 0 and nth[1] > 0 and conics[1, 0] < 0              # det < 0 still rendered
    assert radii[2] > 0 and nth[2] == 0                                  # off-screen: written then culled
    # tiny Gaussian: max(0.1, b^2-det) makes v2 negative -> radius.y = NaN -> `NaN < radius_clip` is false: kept
    assert radii[3] == 2 and nth[3] > 0


def test_tile_row_bands_sum_to_full_gradient(oracle):
    """The multi-GPU tile-row split: per-band partial gradients add up to the full ones."""
    from gaussianimage_plus_b200 import synth

    H, W, N = 96, 128, 300
    xyz, cov, bound, rgb = synth.init_covariance_model(N, H, W, seed=1, colors="rand")
    xys, depths, radii, conics, nth = oracle.project_cov_fwd(xyz, cov + bound, H, W)
    tb = oracle.tile_bounds(H, W)
    total, cum, ids, gids, ids_s, gids_s, bins = oracle.bin_and_sort(xys, depths, radii, nth, tb)
    v_out = np.random.default_rng(0).normal(size=(H, W, 3)).astype(np.float32)
    full = oracle.rasterize_sum_bwd(H, W, gids_s, bins, xys, conics, rgb, None, v_out)
    acc = [np.zeros_like(a) for a in full]
    for band in ((0, 2), (2, 6)):
        b = bins.copy()
        rows = np.arange(tb[0] * tb[1]) // tb[0]
        b[~((rows >= band[0]) & (rows < band[1]))] = 0
        part = oracle.rasterize_sum_bwd(H, W, gids_s, b, xys, conics, rgb, None, v_out)
        acc = [a + p for a, p in zip(acc, part)]
    for a, f in zip(acc, full):
        np.testing.assert_allclose(a, f, rtol=1e-5, atol=1e-6)


def test_ssim_oracle_anchors():
    """oracle/ssim_oracle.py restates pytorch_msssim.ssim (absent, unpinned): its closed-form anchors."""
    import torch

    from oracle import ssim_oracle as S

    win = S.fspecial_gauss_1d()
    assert win.shape == (11,) and abs(float(win.sum()) - 1) < 1e-6 and torch.equal(win, win.flip(0))
    assert abs(float(win[5]) - 0.26601171493530273) < 1e-9      # the float32 value the kernels hard-code
    torch.manual_seed(0)
    x = torch.rand(1, 3, 40, 52, dtype=torch.float64)
    y = torch.rand(1, 3, 40, 52, dtype=torch.float64)
    assert abs(float(S.ssim(x, x)) - 1.0) < 1e-12
    assert abs(float(S.ssim(x, y)) - float(S.ssim(y, x))) < 1e-12
    # constant images: sigma = 0 -> ssim = (2 a b + C1) / (a^2 + b^2 + C1)
    a, b = 0.5, 0.25
    got = float(S.ssim(torch.full((1, 3, 20, 20), a, dtype=torch.float64), torch.full((1, 3, 20, 20), b, dtype=torch.float64)))
    assert abs(got - (2 * a * b + 1e-4) / (a * a + b * b + 1e-4)) < 1e-5
    # valid filtering: an 11x11 image has exactly one window
    assert S.gaussian_filter(x[..., :11, :11], win).shape[-2:] == (1, 1)
    # loss_fn weights (models/utils.py:70-75)
    l2 = float(S.loss_fn(x, y, "L2")); l1 = float(S.loss_fn(x, y, "L1")); ss = float(S.loss_fn(x, y, "SSIM"))
    assert abs(float(S.loss_fn(x, y, "Fusion1")) - (0.7 * l2 + 0.3 * ss)) < 1e-12
    assert abs(float(S.loss_fn(x, y, "Fusion2")) - (0.7 * l1 + 0.3 * ss)) < 1e-12
    assert abs(float(S.loss_fn(x, y, "Fusion3")) - (0.7 * l2 + 0.3 * l1)) < 1e-12


def test_ms_ssim_oracle_anchors():
    import torch

    from oracle import ssim_oracle as S

    torch.manual_seed(0)
    x = torch.rand(1, 3, 176, 200, dtype=torch.float64)
    y = (x + 0.1 * torch.rand_like(x)).clamp(0, 1)
    assert abs(float(S.ms_ssim(x, x)) - 1.0) < 1e-12
    v = float(S.ms_ssim(x, y))
    assert 0.5 < v < 1.0 and abs(v - float(S.ms_ssim(y, x))) < 1e-12
    assert abs(sum(S.MS_WEIGHTS) - 1.0) < 1e-3


def test_ms_ssim_oracle_anchors():
    """oracle/ssim_oracle.ms_ssim (restated pytorch_msssim.ms_ssim; the package is absent: parity unpinned) against
    closed forms: ms_ssim(x, x) == 1 for both window sizes the reference uses (11; 5 in `Fusion_hinerv`), symmetry,
    the window sums to 1, and the two MS-SSIM losses of models/utils.py:76-79 vanish on identical images."""
    import torch

    from oracle import ssim_oracle as S

    torch.manual_seed(0)
    x = torch.rand(1, 3, 176, 200, dtype=torch.float64)
    y = (x + 0.1 * torch.rand_like(x)).clamp(0, 1)
    for win in (11, 5):
        assert abs(float(S.fspecial_gauss_1d(win).sum()) - 1.0) < 1e-6
        assert abs(float(S.ms_ssim(x, x, win_size=win)) - 1.0) < 1e-12
        a, b = float(S.ms_ssim(x, y, win_size=win)), float(S.ms_ssim(y, x, win_size=win))
        assert abs(a - b) < 1e-12 and 0.0 < a < 1.0
    assert float(S.ms_ssim(x, y, win_size=5)) != float(S.ms_ssim(x, y, win_size=11))
    for lt in ("Fusion4", "Fusion_hinerv"):
        assert abs(float(S.loss_fn(x, x, lt))) < 1e-12
        assert float(S.loss_fn(x, y, lt)) > 0


def test_bench_arms_print_the_same_config():
    """bench.py: `--impl reference` runs "on your arm's config" -- both arms build it with one function."""
    import argparse
    import sys

    sys.path.insert(0, ROOT) if "ROOT" in globals() else None
    import bench

    args = argparse.Namespace(workload="kodak_5000", preroll=2000, mode="images", cov_scale=1.0)
    c1, c8 = bench.bench_config(args, 512, 768, 5000, 1), bench.bench_config(args, 512, 768, 5000, 8)
    assert c1["workload"].startswith("kodak_5000: 768x512, 5000 Gaussians") and "l2" in c1 and "flushed" in c1["l2"]
    assert c1 == bench.bench_config(args, 512, 768, 5000, 1) and "tile rows" in c8["parallelism"]
    src = open(bench.__file__).read()
    assert src.count("\"config\": bench_config(args, H, W, N,") == 2      # the GPU arm and the reference arm
