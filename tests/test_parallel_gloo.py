"""CPU, world_size 2, gloo: the host-side multi-GPU logic -- image sharding with a metrics
gather and no data-path collective; tile-row bands whose partial gradients (computed by the oracle on
each rank's band) all-reduce to the single-process gradients."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussianimage_plus_b200.parallel import TileRowPartition, shard_images


def test_shard_images_covers_everything_once():
    for n, w in ((24, 1), (24, 2), (24, 8), (100, 8), (3, 8)):
        got = sorted(i for r in range(w) for i in shard_images(n, w, r))
        assert got == list(range(n))
        sizes = [len(shard_images(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_tile_row_partition():
    for ty, w in ((32, 1), (32, 2), (85, 8), (512, 8), (3, 8)):
        p = TileRowPartition(ty, w)
        bands = [p.band(r) for r in range(w)]
        assert bands[0][0] == 0 and bands[-1][1] == ty
        assert all(bands[i][1] == bands[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in bands]
        assert max(sizes) - min(sizes) <= 1
    load = [1.0] * 10 + [9.0] * 10          # the lower half is 9x heavier
    p = TileRowPartition(20, 2, row_load=load)
    assert p.band(0)[1] > 10                 # the light rows are not enough for half the load
    assert p.owner_of_row(0) == 0 and p.owner_of_row(19) == 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gaussianimage_plus_b200 import synth
    from gaussianimage_plus_b200.parallel import (TileRowPartition, allreduce_packed_gradients, gather_metrics,
                                                  shard_images)
    from oracle import cpu_oracle as O

    # --- mode 1: image set, no collective until the final gather
    mine = shard_images(5, world, rank)
    local = [(i, 20.0 + i, 0.1 * i) for i in mine]
    allm = gather_metrics(local)
    assert [m[0] for m in allm] == list(range(5))
    # --- mode 2: tile-row split of one image
    H, W, N = 96, 128, 300
    xyz, cov, bound, rgb = synth.init_covariance_model(N, H, W, seed=1, colors="rand")
    xys, depths, radii, conics, nth = O.project_cov_fwd(xyz, cov + bound, H, W)
    tb = O.tile_bounds(H, W)
    total, cum, ids, gids, ids_s, gids_s, bins = O.bin_and_sort(xys, depths, radii, nth, tb)
    v_out = np.random.default_rng(0).normal(size=(H, W, 3)).astype(np.float32)
    part = TileRowPartition(tb[1], world)
    r0, r1 = part.band(rank)
    rows = np.arange(tb[0] * tb[1]) // tb[0]
    b = bins.copy()
    b[~((rows >= r0) & (rows < r1))] = 0
    v_xy, v_conic, v_col, _ = O.rasterize_sum_bwd(H, W, gids_s, b, xys, conics, rgb, None, v_out)
    red = allreduce_packed_gradients(torch.from_numpy(v_xy), torch.from_numpy(v_conic), torch.from_numpy(v_col))
    full = O.rasterize_sum_bwd(H, W, gids_s, bins, xys, conics, rgb, None, v_out)
    for a, f in zip(red, full[:3]):
        np.testing.assert_allclose(a.numpy(), f, rtol=1e-5, atol=1e-6)
    # every rank ends with bitwise identical reduced gradients -> replicated Adam stays in lock-step
    mine_bytes = torch.cat([a.reshape(-1) for a in red]).clone()
    other = [torch.zeros_like(mine_bytes) for _ in range(world)]
    dist.all_gather(other, mine_bytes)
    assert all(torch.equal(other[0], o) for o in other)
    # --- global top-k of a per-pixel error map whose rows are spread over the ranks (band-split densification)
    from gaussianimage_plus_b200.parallel import global_topk

    full = torch.from_numpy(np.random.default_rng(5).random(H * W).astype(np.float32))
    rows_lo, rows_hi = r0 * 16 * W, min(H, r1 * 16) * W
    mine_idx = torch.arange(rows_lo, rows_hi, dtype=torch.int64)
    v, i = global_topk((full[rows_lo:rows_hi], mine_idx), 37)
    ref_v, ref_i = torch.topk(full, 37)
    assert torch.equal(v, ref_v) and torch.equal(i, ref_i)
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    from oracle import cpu_oracle

    cpu_oracle.build()
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
