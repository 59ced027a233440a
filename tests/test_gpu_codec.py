"""GPU: the compression pass (train_quantize.py; gaussianimage_plus_b200/codec.py) over the drop-in operators:
quantised forward, quantisation-aware training with the four optimisers, compress / decompress round trip and
the bits-per-pixel accounting of analysis_wo_ec (gaussianimage_covariance.py:469-509)."""
import numpy as np
import pytest
import torch

from gaussianimage_plus_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _psnr(render, gt_hwc):
    return 10 * np.log10(1.0 / float(((render[0].permute(1, 2, 0) - gt_hwc) ** 2).mean()))


@pytest.mark.parametrize("loss_type", ["L2", "Fusion2"])
def test_quantised_training_and_codec_round_trip(loss_type):
    from gaussianimage_plus_b200.codec import QuantizedGaussianImage
    from gaussianimage_plus_b200.fit import GaussianImageFitter

    N, H, W = 1500, 128, 192
    gt = torch.from_numpy(synth.target_image(H, W, seed=5)).to(DEV)
    torch.manual_seed(1)
    fit = GaussianImageFitter(N, H, W, device=DEV, loss_type=loss_type)
    fit.set_target(gt)
    st = fit.fit(600, max_num_points=N, adaptive_add=False)              # warm-up phase (fused step)
    assert st["best_psnr"] > 24
    q = QuantizedGaussianImage.from_fitter(fit)
    assert q.cur_num_points == fit.best_state()["_xyz"].shape[0] and q.loss_type == loss_type
    with torch.no_grad():
        ptq = _psnr(q.forward_quantize()["render"], gt)                   # post-training quantisation
    assert ptq > st["best_psnr"] - 6.0                                    # 12/10/6 bits cost a few dB at most
    first = None
    for i in range(150):
        image, loss, img_loss, vq_loss, psnr = q.train_iter_quantize(gt)
        first = first if first is not None else psnr
        assert np.isfinite(img_loss) and vq_loss == 0
    assert abs(first - ptq) < 1e-3                                        # the first QAT forward IS the PTQ render
    with torch.no_grad():
        qat = _psnr(q.forward_quantize()["render"], gt)
    assert qat > ptq - 0.05, (ptq, qat)                                   # QAT never hurts, usually helps
    # every quantiser parameter received gradient steps
    for name, p in q.named_parameters():
        assert torch.isfinite(p).all(), name
    # codec: compress -> decompress renders what the quantised forward renders (the log quantiser re-ranges per
    # channel on compress, quantize.py:244, so "close", not "equal")
    enc = q.compress_wo_ec()
    n = q.cur_num_points
    assert enc["xyz"].shape == (n, 2) and enc["quant_cholesky_elements"].shape == (n, 3)
    assert enc["feature_dc_index"].shape == (n, 3)
    for key, hi in (("quant_means", 4095), ("quant_cholesky_elements", 1023), ("feature_dc_index", 63)):
        c = enc[key]
        assert torch.equal(c, c.round()) and c.min() >= 0 and c.max() <= hi, key
    dec = _psnr(q.decompress_wo_ec(enc)["render"], gt)
    assert abs(dec - qat) < 0.5, (dec, qat)
    a = q.analysis_wo_ec(enc)
    bits = n * (2 * 12 + 3 * 10 + 3 * 6) + 32 * 2 * 2 + 32 * 3 * 2 + 32 * 3 * 2
    assert abs(a["bpp"] - bits / (H * W)) < 1e-9
    assert abs(a["bpp"] - (a["position_bpp"] + a["cholesky_bpp"] + a["feature_dc_bpp"])) < 1e-12


@pytest.mark.parametrize("loss_type", ["L2", "Fusion2"])
def test_fused_qat_matches_operator_path(loss_type):
    """FusedQuantizedTrainer (quantisers in torch, everything between the de-quantised attributes and their
    gradients in the fused fit step, the iteration replayed from one CUDA graph) against the operator-path
    QuantizedGaussianImage: same loss and gradients for one iteration, same PSNR trajectory over 150."""
    from gaussianimage_plus_b200.codec import (FusedQuantizedTrainer, QuantizedGaussianImage, _FusedRenderLoss,
                                               loss_fn)
    from gaussianimage_plus_b200.fit import GaussianImageFitter

    N, H, W = 1500, 128, 192
    gt = torch.from_numpy(synth.target_image(H, W, seed=5)).to(DEV)
    torch.manual_seed(1)
    fit = GaussianImageFitter(N, H, W, device=DEV, loss_type=loss_type)
    fit.set_target(gt)
    fit.fit(400, max_num_points=N, adaptive_add=False)
    st = fit.best_state()
    args = (st["_xyz"], st["_cov2d"], st["_features_dc"], st["cholesky_bound"], H, W)
    q_op = QuantizedGaussianImage(*args, lr=0.018, loss_type=loss_type)
    q_fu = FusedQuantizedTrainer(*args, lr=0.018, loss_type=loss_type)
    q_fu.set_target(gt)
    # ---- one iteration, gradients only
    loss_op = loss_fn(q_op.forward_quantize()["render"], gt, loss_type, 0.7)
    loss_op.backward()
    means, _, _, _ = q_fu.xyz_quantizer(q_fu._xyz)
    cov, _, _, _ = q_fu.cholesky_quantizer(q_fu.get_cov2d_elements)
    colors, _, _, _ = q_fu.features_dc_quantizer(q_fu.get_features)
    loss_fu = _FusedRenderLoss.apply(means, cov, colors, q_fu)
    loss_fu.backward()
    assert abs(float(loss_op.detach()) - float(loss_fu.detach())) <= 2e-5 * abs(float(loss_op.detach())) + 1e-7
    del loss_op, loss_fu, means, cov, colors          # (nothing may keep this autograd graph alive)
    po, pf = dict(q_op.named_parameters()), dict(q_fu.named_parameters())
    assert set(po) == set(pf)
    for name in po:
        a, b = pf[name].grad.double(), po[name].grad.double()
        rel = float(torch.linalg.norm(a - b) / (torch.linalg.norm(b) + 1e-30))
        assert rel < 5e-4, (name, rel)
    for q in (q_op, q_fu):
        for o in (q.optimizer, q.cov2d_quantizer_optimizer, q.xyz_quantizer_optimizer, q.color_quantizer_optimizer):
            o.zero_grad(set_to_none=True)
    # ---- 150 iterations: eager warm-up + graph replays vs the operator path
    for _ in range(150):
        q_fu.train_iter_quantize()
        _, _, _, _, psnr_op = q_op.train_iter_quantize(gt)
    assert q_fu._graph is not None
    psnr_fu = q_fu.psnr()
    assert abs(psnr_fu - psnr_op) < 0.6, (psnr_fu, psnr_op)
    # the codec half of the parent class works on the fused trainer's parameters
    enc = q_fu.compress_wo_ec()
    assert q_fu.decompress_wo_ec(enc)["render"].shape == (1, 3, H, W)


@pytest.mark.parametrize("color_norm", [False, True])
def test_kernel_qat_matches_operator_path(color_norm):
    """KernelQuantizedTrainer (quantisers, their backward and the four optimisers as kernels, csrc/gi2d_quant.cu)
    against the operator-path QuantizedGaussianImage, whose quantisers are golden-tested against the reference's
    quantize.py (tests/test_quantize_golden.py): quantiser initialisation, de-quantised attributes, every
    gradient of one iteration, the state after one optimiser step, and the PSNR trajectory over 150 iterations."""
    from gaussianimage_plus_b200.codec import KernelQuantizedTrainer, QuantizedGaussianImage, loss_fn
    from gaussianimage_plus_b200.fit import GaussianImageFitter

    N, H, W = 1500, 128, 192
    gt = torch.from_numpy(synth.target_image(H, W, seed=5)).to(DEV)
    torch.manual_seed(1)
    fit = GaussianImageFitter(N, H, W, device=DEV, loss_type="L2", color_norm=color_norm)
    fit.set_target(gt)
    fit.fit(400, max_num_points=N, adaptive_add=False)
    st = fit.best_state()
    args = (st["_xyz"], st["_cov2d"], st["_features_dc"], st["cholesky_bound"], H, W)
    q_op = QuantizedGaussianImage(*args, lr=0.018, loss_type="L2", color_norm=color_norm)
    q_k = KernelQuantizedTrainer(*args, lr=0.018, loss_type="L2", color_norm=color_norm, use_graph=False,
                                 debug_grads=True)
    q_k.set_target(gt)
    # ---- one iteration of both
    means = q_op.xyz_quantizer(q_op._xyz)[0]
    cov = q_op.cholesky_quantizer(q_op.get_cov2d_elements)[0]
    colors = q_op.features_dc_quantizer(q_op.get_features)[0]
    loss_op = loss_fn(q_op._render(means, cov, colors), gt, "L2", 0.7)
    loss_op.backward()
    q_k.train_iter_quantize()
    torch.cuda.synchronize()
    # quantiser initialisation (qparams are updated by the iteration: compare what the module got from the data
    # with a fresh init of a second trainer)
    q_i = KernelQuantizedTrainer(*args, lr=0.018, loss_type="L2", color_norm=color_norm, use_graph=False)
    q_i.init_quantizers()
    qp = q_i.qparams.cpu()
    want = torch.cat([q_op.xyz_quantizer.scale, q_op.xyz_quantizer.beta, q_op.cholesky_quantizer.cov_quantizer.scale,
                      q_op.cholesky_quantizer.cov_quantizer.beta, q_op.features_dc_quantizer.scale,
                      q_op.features_dc_quantizer.beta]).detach().cpu()
    # (the module's parameters have not been stepped yet: optimizer_step() comes below)
    assert torch.allclose(qp, want, rtol=1e-6, atol=1e-9), (qp, want)
    # de-quantised attributes: what the fit step saw
    f = q_k._fit
    for got, ref, name in ((f._t_xyz, means, "means"), (f._t_cov2d, cov, "cov"), (f._t_f_dc, colors, "colors")):
        assert torch.allclose(got[:N], ref.detach(), rtol=2e-6, atol=1e-7), name
    assert abs(q_k.loss() - float(loss_op.detach())) <= 2e-5 * abs(float(loss_op.detach())) + 1e-7
    # gradients of the raw attributes and of the 12 quantiser parameters
    g = q_k.dbg_grads
    for got, ref, name in ((g[:, 0:2], q_op._xyz.grad, "xyz"), (g[:, 2:5], q_op._cov2d.grad, "cov2d"),
                           (g[:, 5:8], q_op._features_dc.grad, "f_dc")):
        rel = float(torch.linalg.norm(got.double() - ref.double()) / (torch.linalg.norm(ref.double()) + 1e-30))
        assert rel < 5e-4, (name, rel)
    qs = q_k.qstats.cpu()
    gs = torch.cat([q_op.xyz_quantizer.scale.grad, q_op.cholesky_quantizer.cov_quantizer.scale.grad,
                    q_op.features_dc_quantizer.scale.grad]).double().cpu()
    gb = torch.cat([q_op.xyz_quantizer.beta.grad, q_op.cholesky_quantizer.cov_quantizer.beta.grad,
                    q_op.features_dc_quantizer.beta.grad]).double().cpu()
    assert float(torch.linalg.norm(qs[7:13] - gs) / (torch.linalg.norm(gs) + 1e-30)) < 2e-3, (qs[7:13], gs)
    assert float(torch.linalg.norm(qs[13:19] - gb) / (torch.linalg.norm(gb) + 1e-30)) < 2e-3, (qs[13:19], gb)
    # the state after one optimiser step (the first Adam step moves every element by ~lr * sign(g): elements whose
    # gradient is within rounding of zero may differ, hence the quantile)
    q_op.optimizer_step()
    for got, ref, name in ((q_k._xyz, q_op._xyz, "xyz"), (q_k._cov2d, q_op._cov2d, "cov2d"),
                           (q_k._features_dc, q_op._features_dc, "f_dc")):
        d = (got.detach() - ref.detach()).abs().flatten()
        assert float(torch.quantile(d, 0.99)) <= 1e-5 * max(1.0, float(ref.detach().abs().max())), name
    want = torch.cat([q_op.xyz_quantizer.scale, q_op.xyz_quantizer.beta, q_op.cholesky_quantizer.cov_quantizer.scale,
                      q_op.cholesky_quantizer.cov_quantizer.beta, q_op.features_dc_quantizer.scale,
                      q_op.features_dc_quantizer.beta]).detach().cpu()
    assert torch.allclose(q_k.qparams.cpu(), want, rtol=1e-4, atol=1e-7), (q_k.qparams.cpu(), want)
    # ---- 150 more iterations, the kernel trainer replaying its CUDA graph
    q_k.use_graph = True
    for _ in range(150):
        q_k.train_iter_quantize()
        _, _, _, _, psnr_op = q_op.train_iter_quantize(gt)
    assert q_k._graph is not None
    psnr_k = q_k.psnr()
    assert abs(psnr_k - psnr_op) < 0.6, (psnr_k, psnr_op)
    assert int(q_k.qstats[0].item()) == 151
    # the codec half of the parent class works on the kernel trainer's state
    enc = q_k.compress_wo_ec()
    dec = _psnr(q_k.decompress_wo_ec(enc)["render"], gt)
    assert abs(dec - psnr_k) < 1.0, (dec, psnr_k)
