"""GPU: the fused fit step (gi2d_fit_forward_backward / gi2d_fit_adam through GaussianImageFitter)
against the oracle's restatement of the reference train_iter, and size-independent properties at
BASELINE.json's full sizes."""
import numpy as np
import pytest
import torch

from gaussianimage_plus_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def N_(t):
    return t.detach().cpu().numpy()


def make_fitter(N, H, W, seed, colors="rand", cov_scale=1.0, use_graph=False, keep_render=True, **kw):
    from gaussianimage_plus_b200.fit import GaussianImageFitter

    xyz, cov, bound, rgb = synth.init_covariance_model(N, H, W, seed=seed, colors=colors, cov_scale=cov_scale)
    gt = synth.target_image(H, W, seed=seed)
    fit = GaussianImageFitter(N, H, W, device=DEV, use_graph=use_graph, **kw)
    fit.keep_render = keep_render
    for dst, src in ((fit._xyz, xyz), (fit._cov2d, cov), (fit.cholesky_bound, bound), (fit._features_dc, rgb)):
        dst.copy_(torch.from_numpy(src))
    fit.set_target(torch.from_numpy(gt))
    fit._bind()
    return fit, (xyz, cov, bound, rgb, gt)


@pytest.mark.parametrize("N,H,W,scale", [(400, 96, 128, 1.0), (5000, 512, 768, 1.0), (5000, 512, 768, 4.0),
                                          (20000, 1356, 2040, 1.0), (3000, 64, 64, 3.0),
                                          (1000000, 8192, 8192, 1.0)])   # BASELINE configs[4], 2^18 tiles
def test_fit_binning_bit_exact(oracle, N, H, W, scale):
    """sorted (tile|gaussian) keys, tile ranges and num_intersects of the fused path == oracle
    (== the reference's cumsum + map + sort + gather + edges), including the multi-pass case
    (2040x1356: 10880 tiles > 2048 bins)."""
    fit, (xyz, cov, bound, rgb, gt) = make_fitter(N, H, W, seed=5, cov_scale=scale)
    fit.forward()
    st = fit.stats()
    tb = oracle.tile_bounds(H, W)
    xys, depths, radii, conics, nth = oracle.project_cov_fwd(xyz, cov + bound, H, W, tb)
    total, cum, ids, gids, ids_s, gids_s, bins = oracle.bin_and_sort(xys, depths, radii, nth, tb)
    assert st["num_intersects"] == total and not st["overflow"]
    keys = N_(fit.sorted_keys[:total])
    np.testing.assert_array_equal(keys >> 32, ids_s >> 32)
    np.testing.assert_array_equal((keys & 0xFFFFFFFF).astype(np.int32), gids_s)
    np.testing.assert_array_equal(N_(fit.tile_bins), bins)
    assert (np.diff(keys) > 0).all()  # a strictly ascending 64-bit sequence
    proj = N_(fit.proj)
    np.testing.assert_array_equal(proj[:, 0:2], xys)
    np.testing.assert_array_equal(proj[:, 2:5], conics)


def test_render_matches_operator_path_bitwise(oracle):
    """fused render == clamp/permute of the stand-alone operator path (same kernels' math, same order)."""
    from gaussianimage_plus_b200.gsplat import project_gaussians_2d_covariance, rasterize_gaussians_plus

    N, H, W = 5000, 512, 768
    fit, (xyz, cov, bound, rgb, gt) = make_fitter(N, H, W, seed=9, cov_scale=2.0)
    render = fit.forward()["render"]
    tb = fit.tile_bounds
    xys, depths, radii, conics, nth = project_gaussians_2d_covariance(fit._xyz, fit._cov2d + fit.cholesky_bound, H, W, tb)
    out = rasterize_gaussians_plus(xys, depths, radii, conics, nth, fit._features_dc, torch.ones(N, 1, device=DEV), H, W)
    expect = torch.clamp(out, 0, 1).view(-1, H, W, 3).permute(0, 3, 1, 2).contiguous()
    assert torch.equal(render, expect)


@pytest.mark.parametrize("N,H,W,scale,graph", [(400, 96, 128, 1.0, False), (2500, 512, 768, 1.0, True),
                                                (5000, 512, 768, 3.0, False)])
def test_fit_step_matches_oracle(oracle, N, H, W, scale, graph):
    fit, (xyz, cov, bound, rgb, gt) = make_fitter(N, H, W, seed=21, cov_scale=scale, use_graph=graph)
    ref = oracle.FitState(xyz, cov, bound, rgb, gt)
    mse_ref, img_ref = ref.train_iter(want_image=True)
    fit.train_iter()
    st = fit.stats()
    assert st["step"] == 1 and st["num_intersects"] == ref.num_intersects
    img = N_(fit.out_hwc)
    bad = np.abs(img - img_ref) > 1e-4 * np.abs(img_ref) + 1e-5
    assert bad.mean() < 1e-5, bad.sum()          # a handful of pixels may sit on the 1/255 threshold
    assert abs(st["mse"] - mse_ref) <= 1e-5 * mse_ref
    assert abs(st["lr"] - 0.018) < 1e-9
    # step 1 of Adam moves every parameter by lr * sign(g): compare where the gradient is not ~0
    for name, a, b0, b in (("xyz", fit._xyz, xyz, ref.xyz), ("cov", fit._cov2d, cov, ref.cov),
                           ("rgb", fit._features_dc, rgb, ref.rgb)):
        d = np.abs(N_(a) - b)
        assert np.quantile(d, 0.995) < 1e-5, (name, np.quantile(d, 0.995))
        moved = np.abs(b - b0) > 1e-3
        assert moved.mean() > 0.3, name


def test_fit_trajectory_tracks_oracle(oracle):
    """PSNR trajectory of the fused path vs the oracle's train_iter restatement at equal iterations."""
    N, H, W = 1500, 128, 192
    fit, (xyz, cov, bound, rgb, gt) = make_fitter(N, H, W, seed=33, colors="zeros", use_graph=True)
    ref = oracle.FitState(xyz, cov, bound, rgb, gt)
    iters = 300
    for _ in range(iters):
        fit.train_iter()
        mse_ref = ref.train_iter()
    st = fit.stats()
    psnr_ref = 10 * np.log10(1.0 / mse_ref)
    assert st["step"] == iters
    assert st["psnr"] > 18.0, st
    assert abs(st["psnr"] - psnr_ref) < 0.25, (st["psnr"], psnr_ref)   # north_star: >= 90% of the trajectory


def test_train_iters_unrolled_graphs_equal_single_steps():
    a, _ = make_fitter(2500, 512, 768, seed=4, colors="zeros", use_graph=True, keep_render=False)
    b, _ = make_fitter(2500, 512, 768, seed=4, colors="zeros", use_graph=True, keep_render=False)
    a.train_iters(37, unroll=8)            # 1 eager + 4 graphs of 8 + 4 single steps
    for _ in range(37):
        b.train_iter()
    sa, sb = a.stats(), b.stats()
    assert sa["step"] == sb["step"] == 37 and sa["num_intersects"] == sb["num_intersects"]
    assert abs(sa["mse"] - sb["mse"]) <= 1e-4 * sb["mse"]
    assert sa["best_step"] == sb["best_step"]
    assert torch.allclose(a._xyz, b._xyz, atol=5e-3)


def test_fit_graph_equals_eager():
    a, _ = make_fitter(2500, 512, 768, seed=4, colors="zeros", use_graph=True)
    b, _ = make_fitter(2500, 512, 768, seed=4, colors="zeros", use_graph=False)
    for _ in range(5):
        a.train_iter()
        b.train_iter()
    sa, sb = a.stats(), b.stats()
    assert sa["step"] == sb["step"] == 5 and sa["num_intersects"] == sb["num_intersects"]
    # float atomics reorder sums: equal to fp32 noise
    assert abs(sa["mse"] - sb["mse"]) <= 1e-4 * sb["mse"]


def test_capacity_overflow_is_detected_and_recovered():
    fit, _ = make_fitter(5000, 512, 768, seed=2, isect_capacity=4096)
    fit.train_iter()
    st = fit.stats()
    assert st["overflow"] and st["num_intersects"] > 4096
    x0 = fit._xyz.clone()
    assert fit.ensure_capacity()
    assert torch.equal(x0, fit._xyz)  # the overflowing step did not touch the parameters
    fit.train_iter()
    st = fit.stats()
    assert not st["overflow"] and st["step"] == 1


def test_fit_loop_grows_the_intersection_buffers_by_itself():
    fit, _ = make_fitter(3000, 256, 384, seed=3, colors="zeros", use_graph=True, keep_render=False, isect_capacity=6000)
    fit.train_iter()
    assert fit.stats()["overflow"]                       # 3000 Gaussians x ~4 tiles do not fit 6000 slots
    st = fit.fit(250, prune_iter=100, adaptive_add=False)
    assert not st["overflow"] and fit.isect_capacity > 6000
    assert st["psnr"] > 15 and st["num_intersects"] > 6000


def test_overflow_steps_are_noops_and_are_rerun():
    """Several iterations pass between an overflow and the host's next look at the flag: every one of
    them must be a no-op on the device -- parameters, step counter, Adam bias correction and StepLR untouched --
    and `catch_up()` / `fit()` must re-run exactly those iterations after growing the buffers."""
    a, _ = make_fitter(3000, 256, 384, seed=3, colors="zeros", use_graph=True, keep_render=False, isect_capacity=6000)
    b, _ = make_fitter(3000, 256, 384, seed=3, colors="zeros", use_graph=True, keep_render=False)
    x0 = a._xyz.clone()
    a.train_iters(12)
    st = a.stats()
    assert st["overflow"] and st["step"] == 0 and st["lr"] == b.stats()["lr"]
    assert torch.equal(x0, a._xyz)
    st = a.catch_up()
    assert st["step"] == 12 and not st["overflow"] and a.isect_capacity > 6000
    b.train_iters(12)
    sb = b.stats()
    assert sb["step"] == 12 and st["num_intersects"] == sb["num_intersects"]
    assert abs(st["mse"] - sb["mse"]) <= 1e-4 * sb["mse"]
    assert torch.allclose(a._xyz, b._xyz, atol=5e-3)
    # an overflow that appears in the middle of a run (the Gaussians grow): fit() without prune checkpoints only
    # looks at the end, and must still deliver every iteration
    i0 = sb["num_intersects"]
    sb = b.fit(288, prune=False, adaptive_add=False)
    assert sb["step"] == 300
    if sb["num_intersects"] > i0 + 16:
        c, _ = make_fitter(3000, 256, 384, seed=3, colors="zeros", use_graph=True, keep_render=False,
                           isect_capacity=(i0 + sb["num_intersects"]) // 2)
        sc = c.fit(300, prune=False, adaptive_add=False)
        assert sc["step"] == 300 and not sc["overflow"] and sc["lr"] == sb["lr"]
        assert abs(sc["psnr"] - sb["psnr"]) < 0.1, (sc["psnr"], sb["psnr"])


def test_prune_and_densify_keep_state_consistent():
    fit, _ = make_fitter(2000, 256, 384, seed=6, colors="zeros", use_graph=True)
    for _ in range(20):
        fit.train_iter()
    p0 = fit.psnr()
    fit._cov2d[:7, 0] = -1000.0  # make 7 Gaussians indefinite
    n_bad, n_now = fit.non_semi_definite_prune()
    assert n_bad == 7 and n_now == 1993 and fit.exp_avg["xyz"].shape[0] == 1993
    added = fit.add_sample_positions(max_num_points=2300, base_num_samples=200)
    # densification_postfix drops the new Gaussians whose random covariance is not positive definite
    assert added == 200 and 2100 < fit.cur_num_points <= 2193
    assert fit.cholesky_bound.shape[0] == fit.cur_num_points == fit.exp_avg_sq['cov2d'].shape[0]
    for _ in range(50):
        fit.train_iter()
    st = fit.stats()
    assert st["step"] == 70 and np.isfinite(st["psnr"]) and st["psnr"] > p0 - 1


def test_device_prune_matches_oracle(oracle):
    """gi2d_fit_prune == the oracle's restatement of non_semi_definite_prune (gaussianimage_covariance.py:336-382):
    the same rows survive, in the same order, with their Adam moments and bounds, bit for bit; the live count and
    the reference's return value; nothing moves when nothing fails or everything would."""
    fit, _ = make_fitter(3000, 256, 384, seed=6, colors="zeros", use_graph=True, keep_render=False)
    fit.train_iters(12)
    g = torch.Generator().manual_seed(5)
    bad = torch.randperm(3000, generator=g)[:137]
    cov = fit._cov2d
    cov[bad[:50], 0] = -5.0                      # sxx < 0
    cov[bad[50:100], 1] = 1e4                    # det < 0
    cov[bad[100:], 2] = -fit.cholesky_bound[bad[100:], 2]   # syy == 0 exactly (singular)
    state = lambda: (N_(fit._xyz), N_(fit._cov2d), N_(fit._features_dc), N_(fit.cholesky_bound),
                     {k: N_(t) for k, t in fit.exp_avg.items()}, {k: N_(t) for k, t in fit.exp_avg_sq.items()})
    before = state()
    n_bad, *want = oracle.non_semi_definite_prune(*before)
    ptrs = (fit._t_xyz.data_ptr(), fit._t_m["cov2d"].data_ptr(), fit.grads.data_ptr())
    got_bad, n_now = fit.non_semi_definite_prune()
    assert (got_bad, n_now) == (n_bad, 3000 - 137) == (137, fit.cur_num_points)
    assert ptrs == (fit._t_xyz.data_ptr(), fit._t_m["cov2d"].data_ptr(), fit.grads.data_ptr())   # nothing reallocated
    after = state()
    for a, b in zip(after[:4], want[:4]):
        np.testing.assert_array_equal(a, b)
    for a, b in zip(after[4:], want[4:]):
        for k in a:
            np.testing.assert_array_equal(a[k], b[k])
    # a second prune finds nothing and moves nothing
    assert fit.non_semi_definite_prune() == (0, 3000 - 137)
    for a, b in zip(state()[:4], want[:4]):
        np.testing.assert_array_equal(a, b)
    # every row invalid: `cur_num_points - to_prune_nums > 0` fails, the model stays
    fit._cov2d[:, 0] = -1000.0
    keep = state()
    assert fit.non_semi_definite_prune() == (3000 - 137, 3000 - 137)
    for a, b in zip(state()[:4], keep[:4]):
        np.testing.assert_array_equal(a, b)
    # the training step keeps working on the compacted model, replayed from the SAME graph
    fit._cov2d.copy_(torch.from_numpy(want[1]))
    fit.train_iters(20)
    st = fit.stats()
    assert st["step"] == 32 and st["num_points"] == 3000 - 137 and np.isfinite(st["psnr"])


def test_device_densify_matches_oracle(oracle):
    """gi2d_fit_densify == the oracle's restatement of add_sample_positions + densification_postfix (train.py:85-118,
    gaussianimage_covariance.py:307-334): the same pixels in the same order, the same rows appended (positions,
    covariances, zero colour / moments, the bound of the NEW count), candidates with a non-positive-definite draw
    skipped, nothing reallocated."""
    H, W = 256, 384
    fit, _ = make_fitter(2000, H, W, seed=6, colors="zeros", use_graph=True, keep_render=False, max_num_points=3000)
    fit.train_iters(30)
    fit.train_iter(want_error_map=True)
    torch.cuda.synchronize()
    errors = N_(fit.err_map)
    before = (N_(fit._xyz), N_(fit._cov2d), N_(fit._features_dc), N_(fit.cholesky_bound))
    ptrs = (fit._t_xyz.data_ptr(), fit._t_bound.data_ptr(), fit._keys_buf.data_ptr())
    torch.manual_seed(77)
    draw = (torch.rand(700, 3) + torch.tensor([0.5, 0, 0.5])).numpy()
    draw[::9, 1] = 3.0                                                    # (make some of them indefinite)
    added = fit.add_sample_positions(max_num_points=3000, base_num_samples=700, errors=fit.err_map,
                                     new_cov2d=torch.from_numpy(draw))
    k, order, valid, new_xyz, new_cov, new_bound = oracle.add_sample_positions(
        errors, 2000, 3000, draw, W, H, base_num_samples=700)
    assert added == k == 700 and 0 < valid.sum() < 700
    n_new = 2000 + int(valid.sum())
    assert fit.cur_num_points == n_new and ptrs == (fit._t_xyz.data_ptr(), fit._t_bound.data_ptr(), fit._keys_buf.data_ptr())
    np.testing.assert_array_equal(N_(fit._xyz)[:2000], before[0])
    np.testing.assert_array_equal(N_(fit._xyz)[2000:], new_xyz)
    np.testing.assert_array_equal(N_(fit._cov2d)[2000:], new_cov)
    np.testing.assert_array_equal(N_(fit.cholesky_bound)[:2000], before[3])
    np.testing.assert_array_equal(N_(fit.cholesky_bound)[2000:], new_bound)
    assert not N_(fit._features_dc)[2000:].any()
    for d in (fit.exp_avg, fit.exp_avg_sq):
        for t in d.values():
            assert not N_(t)[2000:].any()
    # the last growth fills up to max_num_points (train.py:93-94), and the fit goes on from the same graph
    fit.train_iter(want_error_map=True)
    added = fit.add_sample_positions(max_num_points=3000, last=True, errors=fit.err_map)
    assert added == 3000 - n_new and n_new < fit.cur_num_points <= 3000
    fit.train_iters(20)
    st = fit.stats()
    assert st["num_points"] == fit.cur_num_points and np.isfinite(st["psnr"]) and st["step"] == 52


@pytest.mark.parametrize("graph", [False, True])
def test_best_state_snapshot_matches_host_tracking(graph):
    """train.py:132-137 (`if best_psnr < psnr: deepcopy(state_dict)`): the device-side snapshot taken by the
    kernel that applies a step's update == what a host that watched every step would have copied.  The host
    watcher works one step behind and without flushing: after step k+1 was issued the raw parameter tensors
    hold the values right after the update of step k."""
    fit, (_, _, _, _, gt) = make_fitter(1500, 192, 256, seed=21, colors="zeros", use_graph=graph, keep_render=False)
    raw = lambda: torch.cat((fit._t_xyz, fit._t_cov2d, fit._t_f_dc), dim=1).clone()
    best = {"sse": float("inf"), "step": 0, "params": None}
    prev = None          # (sse of step k, k) waiting for the parameters that step k+1's first kernel produces
    fit.lr = 0.5         # (params struct already built: set it there as well) big steps => PSNR goes up AND down
    fit.params.lr0 = 0.5
    fit.reset_stats(0)
    worse = 0
    for it in range(1, 61):
        if it == 41:                           # swap the target: the error jumps, the best stays behind at <= 40
            fit.set_target(torch.from_numpy(1.0 - gt))
        fit.train_iter()
        st = fit.stats()                       # no flush: reads the stats block only
        if prev is not None:
            if prev[0] < best["sse"]:
                best = {"sse": prev[0], "step": prev[1], "params": raw()}
            else:
                worse += 1
        prev = (st["sse"], it)
    hist_ok = worse >= 3                       # the run must have had a non-monotone PSNR to mean anything
    fit.sync_params()                          # flush path: step 60's update (+ its snapshot, if a new best)
    if prev[0] < best["sse"]:
        best = {"sse": prev[0], "step": prev[1], "params": raw()}
    st = fit.stats()
    assert st["best_step"] == best["step"] and st["best_sse"] == best["sse"]
    bs = fit.best_state()
    got = torch.cat((bs["_xyz"], bs["_cov2d"], bs["_features_dc"]), dim=1)
    assert torch.equal(got, best["params"])
    assert hist_ok, "degenerate test: PSNR was monotone"
    assert abs(st["best_psnr"] - 10 * np.log10(3.0 * 192 * 256 / best["sse"])) < 1e-9
    assert st["best_step"] <= 40 < st["step"]


def test_fit_loop_prune_densify_schedule():
    """GaussianImageFitter.fit == train.py:120-152: prune every prune_iter, grow every grow_iter from the L1
    error map of that iteration's render (the last growth fills up to max_num_points), best state kept
    across the changes of the Gaussian count."""
    fit, _ = make_fitter(600, 128, 192, seed=22, colors="zeros", use_graph=True, keep_render=False)
    seen = []
    torch.manual_seed(3047)
    st = fit.fit(900, max_num_points=1500, prune_iter=100, grow_iter=300, callback=lambda it, f: seen.append(f.cur_num_points))
    assert st["step"] == 900
    # +min(1000, room)=900 at 300 (minus non-PSD draws and prunes), then the rest at 600 (= iterations - grow_iter)
    assert seen[298] <= 600 and 600 < seen[299] <= 1500 and seen[-1] <= 1500
    assert seen[599] > seen[598] - 1
    assert fit.cholesky_bound.shape[0] == fit.cur_num_points == fit.exp_avg["xyz"].shape[0]
    # the error map of a step == |clamp(render) - gt| summed over channels, from the same step's render
    fit.keep_render = True
    fit.train_iter(want_error_map=True)
    torch.cuda.synchronize()
    # (keep_render binds out_hwc only through _bind(); bind both for this check)
    fit._bind(out_img=fit.out_hwc.data_ptr(), err_map=True)
    fit._enqueue_step()
    fit._dirty = True
    fit._bind()
    torch.cuda.synchronize()
    ref = (fit.out_hwc.clamp(0, 1) - fit.gt_hwc).abs().sum(dim=2)
    assert torch.allclose(fit.err_map, ref, atol=1e-6)
    bs = fit.best_state()
    assert bs["_xyz"].shape[0] == bs["cholesky_bound"].shape[0]
    assert st["best_psnr"] >= st["psnr"] - 1e-9 and st["best_step"] > 0
    n = fit.load_best_state()
    assert n == bs["_xyz"].shape[0]
    psnr_best_render = 10 * np.log10(1.0 / float(((fit.forward()["render"][0].permute(1, 2, 0) - fit.gt_hwc) ** 2).mean()))
    assert psnr_best_render > st["best_psnr"] - 0.5   # (the snapshot is the state AFTER the best step's update)


def test_overlapped_target_upload_equals_plain():
    """set_target(overlap=True): the next step's target is uploaded into the other of two device buffers on a
    copy stream (bench.py's e2e loop).  Alternating two different images every step must give what the plain,
    in-stream upload gives -- a wrong buffer or a missing dependency shows up as a different error at once."""
    a, (_, _, _, _, gt) = make_fitter(1500, 192, 256, seed=31, colors="zeros", use_graph=True, keep_render=False)
    b, _ = make_fitter(1500, 192, 256, seed=31, colors="zeros", use_graph=True, keep_render=False)
    imgs = [torch.from_numpy(gt).pin_memory(), torch.from_numpy(np.ascontiguousarray(gt[::-1, ::-1])).pin_memory()]
    for i in range(24):
        a.set_target(imgs[i & 1], overlap=True)
        b.set_target(imgs[i & 1])
        a.train_iter()
        b.train_iter()
        if i % 4 == 3 or i < 3:
            sa, sb = a.stats(), b.stats()
            assert sa["step"] == sb["step"] == i + 1
            assert abs(sa["mse"] - sb["mse"]) <= 1e-5 * sb["mse"], (i, sa["mse"], sb["mse"])


def test_step_from_host_equals_train_iter():
    """gi2d_fit_step_host (one C call: pinned target up, step, stats block down) == set_target + train_iter +
    stats(), with two alternating targets so that the double buffering is exercised."""
    from gaussianimage_plus_b200.fit import STAT_COUNT, STAT_STEP

    a, (_, _, _, _, gt) = make_fitter(1500, 192, 256, seed=32, colors="zeros", use_graph=False, keep_render=False)
    b, _ = make_fitter(1500, 192, 256, seed=32, colors="zeros", use_graph=True, keep_render=False)
    u8 = lambda x: torch.from_numpy(np.round(x * 255).astype(np.uint8)).pin_memory()
    imgs = [u8(gt), u8(np.ascontiguousarray(gt[::-1, ::-1]))]
    ring = [torch.zeros(STAT_COUNT, dtype=torch.float64).pin_memory() for _ in range(2)]
    got = []
    prev = None
    for i in range(16):
        slot = a.step_from_host(imgs[i & 1], ring[i & 1])
        if prev is not None:
            a.wait_host_result(prev[0])
            got.append((int(ring[prev[1]][STAT_STEP]), a.mse_from_stats(ring[prev[1]], 192, 256)))
        prev = (slot, i & 1)
    a.wait_host_result(prev[0])
    got.append((int(ring[prev[1]][STAT_STEP]), a.mse_from_stats(ring[prev[1]], 192, 256)))
    for i in range(16):
        b.set_target(imgs[i & 1])
        b.train_iter()
        st = b.stats()
        assert got[i][0] == st["step"] == i + 1
        assert abs(got[i][1] - st["mse"]) <= 1e-5 * st["mse"], (i, got[i], st["mse"])
    assert torch.allclose(a._xyz, b._xyz, atol=2e-3)


def test_checkpoint_round_trip_in_reference_format(tmp_path):
    """save_checkpoint writes the dict train.py:173-175 writes; load_checkpoint (and the reference's own loading
    code, train.py:64-75, which reads 'num_gs', 'gs' and 'slv_bound') restore a model that renders identically."""
    fit, _ = make_fitter(800, 96, 128, seed=41, colors="zeros", use_graph=True, keep_render=False)
    for _ in range(40):
        fit.train_iter()
    path = tmp_path / "gaussian_model.pth.tar"
    fit.save_checkpoint(path)
    ck = torch.load(path, map_location="cpu")
    assert set(ck) == {"gs", "num_gs", "psnr", "ms-ssim", "slv_bound"}
    assert {"_xyz", "_cov2d", "_features_dc", "_opacity", "background", "bound"} <= set(ck["gs"])
    assert ck["num_gs"] == ck["gs"]["_xyz"].shape[0] == ck["slv_bound"].shape[0] == 800
    assert abs(ck["psnr"] - fit.stats()["best_psnr"]) < 1e-9
    fit.load_best_state()
    want = fit.forward()["render"].clone()
    other, _ = make_fitter(300, 96, 128, seed=1, use_graph=True, keep_render=False)   # different size on purpose
    other.load_checkpoint(path)
    assert other.cur_num_points == 800
    assert torch.equal(other.forward()["render"], want)
    other.train_iter()                                   # and training goes on from there
    assert other.stats()["step"] >= 1 and np.isfinite(other.stats()["psnr"])


# --------------------------------------------------------------------------- full-size properties
@pytest.mark.parametrize("name", ["kodak_5000", "div2k_20000"])
def test_full_size_properties(name):
    """Size-independent invariants at BASELINE.json sizes: keys strictly ascending; tile ranges partition
    [0, I) in tile order; per-tile ids ascending; rendering is linear in the colours; the analytic
    colour gradient equals a directional finite difference."""
    H, W, N = synth.CONFIGS[name]
    fit, (xyz, cov, bound, rgb, gt) = make_fitter(N, H, W, seed=12, cov_scale=2.0)
    r1 = fit.forward()["render"].clone()
    st = fit.stats()
    I = st["num_intersects"]
    keys = fit.sorted_keys[:I]
    assert bool((keys[1:] > keys[:-1]).all())
    bins = fit.tile_bins
    cnt = bins[:, 1] - bins[:, 0]
    assert int(cnt.sum()) == I and bool((cnt >= 0).all())
    nz = bins[cnt > 0]
    assert bool((nz[1:, 0] == nz[:-1, 1]).all()) and int(nz[0, 0]) == 0 and int(nz[-1, 1]) == I
    tiles_of_keys = (keys >> 32)
    assert bool((torch.bincount(tiles_of_keys, minlength=bins.shape[0]) == cnt).all())
    # linearity in colour (no clamping active: scale colours down)
    base = fit._features_dc.clone()
    fit._features_dc.copy_(base * 0.01)
    ra = fit.forward()["render"].clone()
    fit._features_dc.copy_(base * 0.02)
    rb = fit.forward()["render"].clone()
    assert torch.allclose(rb, 2 * ra, rtol=1e-5, atol=1e-7)
    # directional derivative of the loss w.r.t. colours vs the analytic gradient of the fused backward
    fit._features_dc.copy_(base * 0.05)
    fit.params.lr0 = 0.0  # freeze parameters: Adam with lr=0
    fit.train_iter()
    g = fit.grads[:, 5:8].clone()
    mse0 = fit.stats()["mse"]
    # probe along the gradient itself: the strongest signal a finite difference can get
    d = g / g.norm() * (base.numel() ** 0.5)
    eps = 2e-3
    fit._features_dc.copy_(base * 0.05 + eps * d)
    fit.train_iter()
    mse1 = fit.stats()["mse"]
    fit._features_dc.copy_(base * 0.05 - eps * d)
    fit.train_iter()
    mse2 = fit.stats()["mse"]
    fd = (mse1 - mse2) / (2 * eps)
    an = float((g.double() * d.double()).sum())
    assert abs(fd - an) <= 4e-2 * abs(an) + 1e-9, (fd, an, mse0)


def test_u8_target_equals_float_target():
    """An 8-bit target handed over as bytes gives bit-identical results to the float image u8/255."""
    N, H, W = 2500, 200, 300
    fa, (xyz, cov, bound, rgb, gt) = make_fitter(N, H, W, seed=8, colors="zeros")
    fb, _ = make_fitter(N, H, W, seed=8, colors="zeros")
    gt_u8 = np.round(gt * 255).astype(np.uint8)
    fa.set_target(torch.from_numpy(gt_u8.astype(np.float32) / np.float32(255)))
    fb.set_target(torch.from_numpy(gt_u8))
    fa.train_iter()
    fb.train_iter()
    assert torch.equal(fa.out_hwc, fb.out_hwc)
    sa, sb = fa.stats()["sse"], fb.stats()["sse"]
    assert abs(sa - sb) <= 1e-12 * sa      # identical per-tile partials, summed by double atomics
    # one more step: the gradients (and so the parameters) agree to float-atomic noise
    fa.train_iter()
    fb.train_iter()
    assert abs(fa.stats()["mse"] - fb.stats()["mse"]) <= 1e-5 * fa.stats()["mse"]


def test_fit_step_with_nothing_on_screen_and_tiny_images():
    """Edge cases of the fit step: no Gaussian intersects the image (the reference returns ones * background and
    no gradient, rasterize_sum_plus.py:110-118), an image smaller than one tile, zero-size tails."""
    from gaussianimage_plus_b200.fit import GaussianImageFitter

    for H, W, N in ((40, 56, 50), (7, 9, 12), (16, 16, 1)):
        fit = GaussianImageFitter(N, H, W, device=DEV, use_graph=False)
        fit.set_target(torch.full((H, W, 3), 0.25, device=DEV))
        fit._xyz[:] = torch.tensor([-500.0, -500.0], device=DEV)        # everything far off screen
        x0, c0 = fit._xyz.clone(), fit._features_dc.clone()
        for _ in range(3):
            fit.train_iter()
        st = fit.stats()
        assert st["num_intersects"] == 0 and st["step"] == 3 and not st["overflow"]
        assert abs(st["mse"] - 0.75 ** 2) < 1e-6                        # render == 1 everywhere
        assert torch.equal(fit._xyz, x0) and torch.equal(fit._features_dc, c0)   # zero gradient: Adam does not move
        assert float((fit.forward()["render"] - 1.0).abs().max()) == 0.0
        # and back on screen it fits
        fit._xyz[:] = torch.rand(N, 2, device=DEV) * torch.tensor([W, H], device=DEV)
        for _ in range(30):
            fit.train_iter()
        st2 = fit.stats()
        assert st2["num_intersects"] > 0 and np.isfinite(st2["psnr"]) and st2["mse"] < 0.75 ** 2


def test_tile_row_load_drives_the_band_partition():
    from gaussianimage_plus_b200.parallel import TileRowPartition

    fit, _ = make_fitter(3000, 256, 384, seed=9, colors="zeros", use_graph=False, keep_render=False)
    fit._xyz[:, 1] = fit._xyz[:, 1] * 0.5           # every Gaussian in the upper half of the image
    fit.train_iter()
    load = fit.tile_row_load()
    assert load.shape == (16,) and float(load.sum()) == fit.stats()["num_intersects"]
    assert float(load[:8].sum()) > 0.95 * float(load.sum())
    even = TileRowPartition(16, 2)
    part = TileRowPartition(16, 2, row_load=load.tolist())
    assert even.band(0) == (0, 8) and part.band(0)[1] < 8      # the loaded half is split between the ranks
    halves = [float(load[a:b].sum()) for a, b in (part.band(0), part.band(1))]
    assert max(halves) / sum(halves) < 0.75


def test_fit_with_color_norm_matches_oracle(oracle):
    """color_norm (colours = sigmoid(features), gaussianimage_covariance.py:74,160-162; the setting of the
    compression pass): several fused steps against the oracle's train_iter with the same activation."""
    N, H, W = 1200, 96, 144
    fit, (xyz, cov, bound, rgb, gt) = make_fitter(N, H, W, seed=13, colors="rand", cov_scale=1.5, use_graph=True,
                                                  color_norm=True)
    ref = oracle.FitState(xyz, cov, bound, rgb, gt, color_sigmoid=True)
    for it in range(6):
        mse_ref, img_ref = ref.train_iter(want_image=True)
        fit.train_iter()
        st = fit.stats()
        assert st["step"] == it + 1 and st["num_intersects"] == ref.num_intersects
        assert abs(st["mse"] - mse_ref) <= 2e-4 * mse_ref, (it, st["mse"], mse_ref)
    for a, b in ((fit._xyz, ref.xyz), (fit._cov2d, ref.cov), (fit._features_dc, ref.rgb)):
        d = np.abs(N_(a) - b)
        assert np.quantile(d, 0.99) < 2e-3, np.quantile(d, 0.99)
    r = N_(fit.forward()["render"])
    assert r.min() >= 0 and r.max() <= 1


@pytest.mark.parametrize("N", [700, 1500])   # the fullest tile: ranked in shared memory (<= 960) / radix-selected (beyond)
def test_bucket_overflow_of_one_tile_is_vetoed_and_regrown(oracle, N):
    """Bucketed binning: tile t owns capacity / #tiles rows.  A scene whose Gaussians all sit
    in ONE tile overflows that bucket long before the total count reaches the capacity: the step must be an
    optimiser no-op (parameters, step counter untouched), GI2D_STAT_MAX_TILE must tell the host how far to regrow,
    and after ONE regrow the binning must equal the oracle's bit for bit (incl. the >256-entries-per-tile path)."""
    from gaussianimage_plus_b200.fit import GaussianImageFitter

    H, W = 128, 192                                          # 96 tiles
    rng = np.random.default_rng(9)
    xyz = (np.array([40.0, 40.0], np.float32) + rng.uniform(-3, 3, (N, 2))).astype(np.float32)   # all inside tile (2,2)
    cov = np.tile(np.array([1.0, 0.0, 1.0], np.float32), (N, 1)) + rng.uniform(0, 0.2, (N, 3)).astype(np.float32)
    cov[:, 1] = 0.0
    bound = np.zeros((N, 3), np.float32)
    rgb = rng.uniform(0, 0.01, (N, 3)).astype(np.float32)
    gt = synth.target_image(H, W, seed=9)
    fit = GaussianImageFitter(N, H, W, device=DEV, use_graph=False, isect_capacity=96 * 64)   # 64 rows per tile
    assert fit.bucket_cap == 64
    for dst, src in ((fit._xyz, xyz), (fit._cov2d, cov), (fit.cholesky_bound, bound), (fit._features_dc, rgb)):
        dst.copy_(torch.from_numpy(src))
    fit.set_target(torch.from_numpy(gt))
    fit._bind()
    x0 = fit._xyz.clone()
    fit.train_iter()
    fit.train_iter()
    st = fit.stats()
    tb = oracle.tile_bounds(H, W)
    xys, depths, radii, conics, nth = oracle.project_cov_fwd(xyz, cov + bound, H, W, tb)
    total, cum, ids, gids, ids_s, gids_s, bins = oracle.bin_and_sort(xys, depths, radii, nth, tb)
    per_tile = np.bincount((ids_s >> 32).astype(np.int64), minlength=tb[0] * tb[1])
    assert per_tile.max() > 256 and total < fit.isect_capacity          # one bucket overflows, the total would fit
    assert st["overflow"] and st["step"] == 0 and st["num_intersects"] == total
    assert st["max_tile"] == per_tile.max()
    fit.sync_params()
    assert torch.equal(x0, fit._xyz)                                     # the vetoed steps did not touch anything
    st = fit.catch_up(st)                                                # regrow (from max_tile) + re-run the two lost iterations
    assert st["step"] == 2 and not st["overflow"]
    assert fit.bucket_cap >= per_tile.max()
    # the binning of a fresh forward of the START parameters == the oracle's, bit for bit
    fit2 = GaussianImageFitter(N, H, W, device=DEV, use_graph=False, isect_capacity=fit.isect_capacity)
    for dst, src in ((fit2._xyz, xyz), (fit2._cov2d, cov), (fit2.cholesky_bound, bound), (fit2._features_dc, rgb)):
        dst.copy_(torch.from_numpy(src))
    fit2.set_target(torch.from_numpy(gt))
    fit2.keep_render = True
    fit2._bind()
    out = fit2.forward()["render"]
    st2 = fit2.stats()
    assert not st2["overflow"] and st2["num_intersects"] == total
    keys = N_(fit2.sorted_keys[:total])
    np.testing.assert_array_equal(keys >> 32, ids_s >> 32)
    np.testing.assert_array_equal((keys & 0xFFFFFFFF).astype(np.int32), gids_s)
    np.testing.assert_array_equal(N_(fit2.tile_bins), bins)
    # ... and the render of the over-full tile is the reference's: the 256 smallest ids of the sorted list
    img_ref = oracle.rasterize_sum_fwd(H, W, gids_s, bins, xys, conics, rgb)[0]
    got = N_(out[0].permute(1, 2, 0))
    assert np.allclose(got, np.clip(img_ref, 0, 1), rtol=1e-4, atol=1e-6), np.abs(got - np.clip(img_ref, 0, 1)).max()
