"""GPU, world_size 2 (needs two devices; skipped on a one-GPU box): the tile-row split of ONE image
 with both exchange steps --

  * `TileRowPartition.make_grad_hook()`: NCCL all-reduce of the packed gradients + replicated Adam;
  * `FusedTileRowExchange`: gi2d_fit_exchange_adam, ONE kernel doing reduce-scatter (P2P loads) +
    projection backward + Adam on the owned slice + all-gather (P2P stores) over NVLink peer memory --

against the single-GPU fit of the whole image: same parameters (up to the order of the float sums),
same PSNR, and parameters bitwise identical on every rank after every exchange.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

H, W, N, STEPS = 256, 384, 3000, 25


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys

    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gaussianimage_plus_b200 import synth
    from gaussianimage_plus_b200.fit import GaussianImageFitter
    from gaussianimage_plus_b200.parallel import FusedTileRowExchange, TileRowPartition

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NCCL_DEBUG="WARN")
    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    xyz, cov, bound, rgb = synth.init_covariance_model(N, H, W, seed=11, colors="zeros")
    gt = torch.from_numpy(synth.target_image(H, W, seed=11))

    def make(tile_rows=None, hook=None):
        fit = GaussianImageFitter(N, H, W, device=dev, use_graph=False, tile_rows=tile_rows, grad_hook=hook)
        for dst, src in ((fit._xyz, xyz), (fit._cov2d, cov), (fit.cholesky_bound, bound), (fit._features_dc, rgb)):
            dst.copy_(torch.from_numpy(src))
        fit.set_target(gt)
        return fit

    part = TileRowPartition((H + 15) // 16, world)
    whole = make()
    nccl = make(part.band(rank), part.make_grad_hook())
    fused = make(part.band(rank))
    ex = FusedTileRowExchange(fused)
    for _ in range(STEPS):
        whole.train_iter()
        nccl.train_iter()
        fused.train_iter()
    torch.cuda.synchronize(dev)
    ref = {k: t.clone() for k, t in (("xyz", whole._xyz), ("cov", whole._cov2d), ("rgb", whole._features_dc))}
    res = {}
    for name, fit in (("nccl", nccl), ("fused", fused)):
        got = {"xyz": fit._xyz, "cov": fit._cov2d, "rgb": fit._features_dc}
        for k in ref:
            # replicated state must be bitwise identical on every rank
            mine = got[k].contiguous().view(torch.int32).clone()
            parts = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            assert all(torch.equal(parts[0], q) for q in parts), (name, k, "ranks diverged")
            # Adam with eps=1e-15 turns a near-cancelling gradient sum into a +-lr step whose sign is decided by
            # the order of the fp32 additions: judge the bulk (99.9 %) of the entries, bound the rest loosely
            d = (got[k] - ref[k]).abs().flatten()
            res[(name, k)] = float(torch.quantile(d, 0.999))
            assert float(d.max()) < STEPS * 0.018 * 1.01, (name, k, float(d.max()))
    st_whole = whole.stats()
    st_fused = ex.global_stats()
    st_nccl = nccl.stats()
    assert st_whole["step"] == st_fused["step"] == st_nccl["step"] == STEPS
    if rank == 0:
        np.save(os.path.join(out_dir, "res.npy"),
                np.array([res[("nccl", "xyz")], res[("nccl", "cov")], res[("nccl", "rgb")],
                          res[("fused", "xyz")], res[("fused", "cov")], res[("fused", "rgb")],
                          st_whole["psnr"], st_nccl["psnr"], st_fused["psnr"]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_tilerow_exchanges_match_single_gpu(tmp_path):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r = np.load(tmp_path / "res.npy")
    # 25 Adam steps of lr 0.018: each step moves a parameter by <= lr; the two exchanges differ from the
    # single-GPU run only in the order of the fp32 gradient sums (atomics / ring / rank order)
    assert r[0:6].max() < 2e-2, r
    assert abs(r[6] - r[7]) < 0.02 and abs(r[6] - r[8]) < 0.02, r
