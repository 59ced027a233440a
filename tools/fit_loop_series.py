import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from gaussianimage_plus_b200 import synth
from gaussianimage_plus_b200.fit import GaussianImageFitter
H, W, _ = synth.CONFIGS["kodak_5000"]
xyz, cov, bound, rgb = synth.init_covariance_model(2500, H, W, seed=3047, colors="zeros")
gt_u8 = torch.from_numpy(np.round(synth.target_image(H, W) * 255.0).astype(np.uint8))
fit = GaussianImageFitter(2500, H, W, device="cuda:0", max_num_points=5000)
for dst, src in ((fit._xyz, xyz), (fit._cov2d, cov), (fit.cholesky_bound, bound), (fit._features_dc, rgb)):
    dst.copy_(torch.from_numpy(src))
fit.set_target(gt_u8)
fit.train_iters(20); torch.cuda.synchronize()
ev=[torch.cuda.Event(enable_timing=True) for _ in range(60)]
rows=[]
it=0
for k in range(50):
    ev[k].record()
    fit.train_iters(99); fit.train_iter(want_error_map=((it+100)%1000==0 and it+100<5000))
    it+=100
    ev[k+1].record()
    torch.cuda.synchronize()
    t_step=ev[k].elapsed_time(ev[k+1])*10
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); fit.prune_async(); e1.record(); torch.cuda.synchronize(); t_prune=e0.elapsed_time(e1)*1e3
    t_grow=0
    if it%1000==0 and it<5000:
        e0.record(); fit.add_sample_positions(5000, last=(it==4000), errors=fit.err_map); e1.record(); torch.cuda.synchronize(); t_grow=e0.elapsed_time(e1)*1e3
    st=fit.stats()
    rows.append((it, round(t_step,1), round(t_prune), round(t_grow), st['num_points'], st['num_intersects'], st['overflow'], round(st['psnr'],2)))
for r in rows[::3]: print(r)
bins = fit.tile_bins
cnt = (bins[:, 1] - bins[:, 0]).clamp(min=0)
top = torch.sort(cnt, descending=True).values[:12].tolist()
print("largest tile counts at the end:", top, "mean", float(cnt.float().mean()))
import bench, ctypes as C
from gaussianimage_plus_b200 import _lib
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
rec = bench.roofline_record(torch, lib, _lib, fit, flush, "kodak_5000", 74.4)
print({k: rec[k] for k in ("step_ms", "kernel_ms_back_to_back_l2_warm", "step_kernel_ms", "pairs_per_launch", "num_intersects")})
