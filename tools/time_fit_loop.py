#!/usr/bin/env python
"""tools/time_fit_loop.py [repeats]: wall clock of configs[1] as written (bench.bench_fit_loop's loop), one line per
fit with what made it slow when it was: the largest tile of the final state, whether the intersection buffers had
to be regrown.  Not a bench value."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gaussianimage_plus_b200 import synth
from gaussianimage_plus_b200.fit import GaussianImageFitter

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H, W, _ = synth.CONFIGS["kodak_5000"]
xyz, cov, bound, rgb = synth.init_covariance_model(2500, H, W, seed=3047, colors="zeros")
gt_u8 = torch.from_numpy(np.round(synth.target_image(H, W) * 255.0).astype(np.uint8))
for trial in range(reps):
    fit = GaussianImageFitter(2500, H, W, device="cuda:0", max_num_points=5000)
    for dst, src in ((fit._xyz, xyz), (fit._cov2d, cov), (fit.cholesky_bound, bound), (fit._features_dc, rgb)):
        dst.copy_(torch.from_numpy(src))
    fit.set_target(gt_u8)
    cap0 = fit.isect_capacity
    events = []
    _ens, _grow = fit.ensure_capacity, fit.add_sample_positions

    def ens(st=None):
        r = _ens(st)
        if r:
            events.append(("regrow", fit.isect_capacity))
        return r

    def grow(*a, **k):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = _grow(*a, **k)
        torch.cuda.synchronize()
        events.append(("densify_ms", round((time.perf_counter() - t) * 1e3, 2), "queue_s_before", round(t - t0, 3)))
        return r

    fit.ensure_capacity, fit.add_sample_positions = ens, grow
    _ti = fit.train_iters
    seg = {"host_s": 0.0, "calls": 0, "slowest_call_s": 0.0}

    def ti(n, *a, **k):
        t = time.perf_counter()
        r = _ti(n, *a, **k)
        d = time.perf_counter() - t
        seg["host_s"] += d
        seg["calls"] += 1
        seg["slowest_call_s"] = max(seg["slowest_call_s"], d)
        return r

    fit.train_iters = ti
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_a.record()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = fit.fit(5000, max_num_points=5000, prune_iter=100, grow_iter=1000)
    ev_b.record()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    seg = {k: round(v, 4) if isinstance(v, float) else v for k, v in seg.items()}
    seg["gpu_span_s"] = round(ev_a.elapsed_time(ev_b) * 1e-3, 4)
    events.append(seg)
    bins = fit.tile_bins
    cnt = (bins[:, 1] - bins[:, 0]).clamp(min=0)
    print(json.dumps({"trial": trial, "it_s": round(5000 / dt), "seconds": round(dt, 4),
                      "largest_tiles": torch.sort(cnt, descending=True).values[:2].tolist(),
                      "regrown": fit.isect_capacity != cap0, "bucket_cap": fit.bucket_cap, "events": events}))
    del fit
