#!/usr/bin/env python
"""tools/cull_stats.py [workload] [iterations]: CPU-side count of the pixel x Gaussian pairs each culling scheme of
the rasterizer evaluates, on a scene fitted by the oracle's C port (no GPU needed).  Schemes: nominal (every pair of
a tile's list), the forward's 8x8 quadrant masks from the reach box / from the exact ellipse, the backward's
union-of-four-rows sweep (round-2 first version) and its per-Gaussian shaped sweep, and the ideal (pairs with
alpha >= 1/255).  Used to decide where the rasterizer's instructions go; prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from gaussianimage_plus_b200 import synth
from oracle import cpu_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "kodak_5000"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
H, W, N = synth.CONFIGS[name]
xyz, cov, bound, rgb = synth.init_covariance_model(N, H, W, seed=3047, colors="zeros")
gt = np.round(synth.target_image(H, W) * 255.0).astype(np.uint8).astype(np.float32) / 255.0
st = O.FitState(xyz, cov, bound, rgb, gt)
for _ in range(iters):
    st.train_iter()
tb = O.tile_bounds(H, W)
xys, depths, radii, conics, nth = O.project_cov_fwd(st.xyz, st.cov + bound, H, W, tb)
total, _, _, _, ids_s, gids_s, bins = O.bin_and_sort(xys, depths, radii, nth, tb)
tiles_x = tb[0]
L = 5.5412635 * 1.001 + 1e-3
tile_of = (ids_s >> 32).astype(np.int64)
g = gids_s.astype(np.int64)
tx0 = (tile_of % tiles_x) * 16.0
ty0 = (tile_of // tiles_x) * 16.0
gx = xys[g, 0] - tx0
gy = xys[g, 1] - ty0
a, b, c = conics[g, 0], conics[g, 1], conics[g, 2]
det = a * c - b * b
ok = (det > 1e-3 * a * c) & (a > 0) & (c > 0)
k = 2 * L / np.where(ok, det, 1)
hx = np.sqrt(k * c) * 1.001 + 1e-3
hy = np.sqrt(k * a) * 1.001 + 1e-3
x0, x1, y0, y1 = gx - hx, gx + hx, gy - hy, gy + hy
x0 = np.where(ok, x0, -1e9); x1 = np.where(ok, x1, 1e9); y0 = np.where(ok, y0, -1e9); y1 = np.where(ok, y1, 1e9)
I = len(g)
nominal = I * 256
# per-pixel exact acceptance
px = np.arange(16, dtype=np.float32)
dx = gx[:, None] - px[None, :]            # [I,16]
dy = gy[:, None] - px[None, :]
sig = 0.5 * (a[:, None, None] * dx[:, None, :] ** 2 + c[:, None, None] * dy[:, :, None] ** 2) + b[:, None, None] * dx[:, None, :] * dy[:, :, None]
acc = (sig >= 0) & (np.exp(-sig) >= 1 / 255.0)          # [I,16(y),16(x)]
ideal = int(acc.sum())
# quadrant masks from the box
quad_box = np.zeros((I, 2, 2), bool)
quad_exact = np.zeros((I, 2, 2), bool)
for qr in range(2):
    for qc in range(2):
        quad_box[:, qr, qc] = (x0 <= 8 * qc + 7) & (x1 >= 8 * qc) & (y0 <= 8 * qr + 7) & (y1 >= 8 * qr)
        quad_exact[:, qr, qc] = acc[:, 8 * qr:8 * qr + 8, 8 * qc:8 * qc + 8].any(axis=(1, 2))
fwd_box = int(quad_box.sum()) * 64
fwd_exact = int(quad_exact.sum()) * 64
# 4x4 cells (what a finer forward could reach)
cell_exact = sum(int(acc[:, 4 * r:4 * r + 4, 4 * q:4 * q + 4].any(axis=(1, 2)).sum()) for r in range(4) for q in range(4)) * 16
row_lo = np.clip(np.ceil(y0), 0, 15).astype(int); row_hi = np.clip(np.floor(y1), 0, 15).astype(int)
col_lo = np.clip(np.ceil(x0), 0, 15).astype(int); col_hi = np.clip(np.floor(x1), 0, 15).astype(int)
hit = quad_box.any(axis=(1, 2))
# backward, old: groups of four consecutive list entries of a tile, union of rows x 16 columns
# backward, new: shapes + 4 buckets
old_pairs = 0
new_pairs = 0
new_sorted_exact = 0
order = np.argsort(tile_of, kind="stable")
starts = np.flatnonzero(np.r_[True, tile_of[order][1:] != tile_of[order][:-1]])
ends = np.r_[starts[1:], I]
c0 = col_lo & ~1
wd = col_hi - c0
lg = np.where(wd < 4, 1, np.where(wd < 8, 2, 3))
rpt = 8 >> lg
trips = np.maximum(0, (row_hi - row_lo + rpt) // rpt)
for s, e in zip(starts, ends):
    idx = order[s:e]
    idx = idx[hit[idx]][:256]
    if len(idx) == 0:
        continue
    # old
    for q in range(0, len(idx), 4):
        grp = idx[q:q + 4]
        old_pairs += (row_hi[grp].max() - row_lo[grp].min() + 1) * 16 * 4
    t = trips[idx]
    bucket = np.where(t > 8, 0, np.where(t > 4, 1, np.where(t > 2, 2, 3)))
    o = np.argsort(bucket, kind="stable")
    ts = t[o]
    for q in range(0, len(ts), 4):
        new_pairs += ts[q:q + 4].max() * 16 * 4
    ts2 = np.sort(t)[::-1]
    for q in range(0, len(ts2), 4):
        new_sorted_exact += ts2[q:q + 4].max() * 16 * 4
print(json.dumps({"workload": name, "iterations": iters, "num_intersects": int(total), "nominal_pairs": nominal,
                  "ideal_alpha_pairs": ideal, "fwd_quadrant_box": fwd_box, "fwd_quadrant_exact": fwd_exact,
                  "fwd_cell4x4_exact": cell_exact, "bwd_union_rows_x16": int(old_pairs),
                  "bwd_shaped_4buckets": int(new_pairs), "bwd_shaped_exact_sort": int(new_sorted_exact),
                  "fractions_of_nominal": {k: round(v / nominal, 3) for k, v in {
                      "ideal": ideal, "fwd_box": fwd_box, "fwd_exact": fwd_exact, "fwd_4x4": cell_exact,
                      "bwd_old": old_pairs, "bwd_new": new_pairs, "bwd_new_exact_sort": new_sorted_exact}.items()}}))
