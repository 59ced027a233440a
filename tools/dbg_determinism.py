import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianimage_plus_b200 import synth
from gaussianimage_plus_b200.fit import GaussianImageFitter

def make(graph):
    N, H, W = 2500, 512, 768
    xyz, cov, bound, rgb = synth.init_covariance_model(N, H, W, seed=4, colors="zeros")
    gt = synth.target_image(H, W, seed=4)
    fit = GaussianImageFitter(N, H, W, use_graph=graph)
    for dst, src in ((fit._xyz, xyz), (fit._cov2d, cov), (fit.cholesky_bound, bound), (fit._features_dc, rgb)):
        dst.copy_(torch.from_numpy(src))
    fit.set_target(torch.from_numpy(gt))
    return fit

fits = {"g1": make(True), "g2": make(True), "e1": make(False), "e2": make(False)}
for step in range(6):
    row = []
    for k, f in fits.items():
        f.train_iter()
        torch.cuda.synchronize()
        st = f.stats()
        row.append(f"{k}: mse={st['mse']:.8f} I={st['num_intersects']}")
    print(step + 1, " | ".join(row))
# keys sorted?
for k, f in fits.items():
    n = f.stats()["num_intersects"]
    keys = f.sorted_keys[:n]
    print(k, "sorted:", bool((keys[1:] > keys[:-1]).all()))
