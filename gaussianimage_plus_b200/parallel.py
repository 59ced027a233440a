"""Multi-GPU partitioning of the hot path.  One process per GPU, torch.distributed for
the plumbing; the reference itself is single-GPU (`train.py:39`), so both modes are new design.

1. Image sets (Kodak-24, DIV2K-100): images are independent units -> `shard_images` deals image i to
   rank i mod world; NO collective on the data path; `gather_metrics` collects the per-image scalars
   once at the end (the averages of train.py:327-340).

2. One very large image: `TileRowPartition` gives rank r a contiguous band of tile rows.  Parameters
   and Adam state are replicated; each rank projects every Gaussian, bins/rasterizes only its band
   (gi2d_fit_params.tile_row_begin/end) and accumulates partial per-Gaussian gradients; ONE
   all-reduce(SUM) of the packed f32[N,8] gradient buffer (+ the 64 squared-error partials) per
   iteration makes them identical everywhere, after which the replicated Adam step is bitwise the
   same on every rank.  The all-reduce runs on the stream the kernels run on, directly on the buffer
   the backward kernel accumulated into (no staging copy).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_images(num_images: int, world_size: int, rank: int) -> List[int]:
    """Indices of the images rank `rank` fits (round-robin keeps 768x512 / 512x768 mixes balanced)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, num_images, world_size))


def gather_metrics(local: Sequence[Tuple[int, float, float]], group=None):
    """All ranks contribute [(image_index, psnr, seconds), ...]; every rank gets the full sorted list."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return sorted(local)
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, list(local), group=group)
    return sorted(x for part in out for x in part)


class TileRowPartition:
    """Contiguous, near-equal bands of tile rows; optionally weighted by a per-row load histogram
    (adaptive densification concentrates Gaussians, SURVEY 8e 'load balance caveat')."""

    def __init__(self, tiles_y: int, world_size: int, row_load: Sequence[float] = None):
        if world_size < 1 or tiles_y < 0:
            raise ValueError("bad partition")
        self.tiles_y, self.world_size = tiles_y, world_size
        if row_load is None:
            base, rem = divmod(tiles_y, world_size)
            sizes = [base + (1 if r < rem else 0) for r in range(world_size)]
            edges = [0]
            for s in sizes:
                edges.append(edges[-1] + s)
        else:
            if len(row_load) != tiles_y:
                raise ValueError("row_load must have one entry per tile row")
            total = float(sum(row_load)) or 1.0
            edges, acc, r = [0], 0.0, 1
            for y, w in enumerate(row_load):
                acc += w
                while r < world_size and acc >= total * r / world_size:
                    edges.append(y + 1)
                    r += 1
            while len(edges) < world_size:
                edges.append(tiles_y)
            edges.append(tiles_y)
            edges = [min(e, tiles_y) for e in edges]
        self.edges = edges

    def band(self, rank: int) -> Tuple[int, int]:
        return self.edges[rank], self.edges[rank + 1]

    def owner_of_row(self, tile_row: int) -> int:
        for r in range(self.world_size):
            if self.edges[r] <= tile_row < self.edges[r + 1]:
                return r
        raise ValueError(tile_row)

    def make_grad_hook(self, group=None):
        """Hook for GaussianImageFitter(grad_hook=...): all-reduce the packed gradients and the SSE."""
        import torch.distributed as dist

        from .fit import STAT_SSE, STAT_SSE_SLOTS

        def hook(fit):
            dist.all_reduce(fit.grads, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(fit.stats_buf[STAT_SSE:STAT_SSE + STAT_SSE_SLOTS], op=dist.ReduceOp.SUM, group=group)

        return hook


def allreduce_packed_gradients(v_xy, v_conic, v_colors, group=None):
    """Host-side reference of the exchange step for the operator path: pack [N,2]+[N,3]+[N,3] into one
    [N,8] buffer, one all-reduce, unpack.  (The fused path accumulates straight into the packed buffer.)"""
    import torch
    import torch.distributed as dist

    packed = torch.cat((v_xy, v_conic, v_colors), dim=1).contiguous()
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed[:, 0:2], packed[:, 2:5], packed[:, 5:8]


def global_topk(values, k: int, group=None):
    """Top-k over a tensor whose entries are spread over the ranks (densification of a band-split
    image needs the k pixels of largest error of the WHOLE image): every rank contributes its local top-k
    candidates (value, global index), one all-gather of 2k numbers per rank, re-selection everywhere.
    `values`: (local_values f32[n], global_index i64[n]).  Returns (values[k'], global_index[k']) sorted by
    descending value, identical on every rank (ties broken by the smaller global index); k' = min(k, total)."""
    import torch
    import torch.distributed as dist

    vals, idx = values
    kk = min(int(k), vals.numel())
    top_v, pos = torch.topk(vals, kk) if kk else (vals[:0], idx[:0])
    top_i = idx[pos] if kk else idx[:0]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        pad_v = torch.full((int(k),), float("-inf"), dtype=vals.dtype, device=vals.device)
        pad_i = torch.full((int(k),), -1, dtype=torch.int64, device=vals.device)
        pad_v[:kk], pad_i[:kk] = top_v, top_i
        all_v = [torch.empty_like(pad_v) for _ in range(world)]
        all_i = [torch.empty_like(pad_i) for _ in range(world)]
        dist.all_gather(all_v, pad_v, group=group)
        dist.all_gather(all_i, pad_i, group=group)
        top_v, top_i = torch.cat(all_v), torch.cat(all_i)
        keep = top_i >= 0
        top_v, top_i = top_v[keep], top_i[keep]
    # deterministic order: value descending, then index ascending
    order = torch.argsort(top_i)
    top_v, top_i = top_v[order], top_i[order]
    order = torch.argsort(top_v, descending=True, stable=True)
    kk = min(int(k), top_v.numel())
    return top_v[order][:kk], top_i[order][:kk]


class FusedTileRowExchange:
    """Tile-row split with the exchange step fused into ONE kernel over NVLink peer memory.

    Replaces `TileRowPartition.make_grad_hook()` (NCCL all-reduce of grads + replicated Adam) by
    `gi2d_fit_exchange_adam`: reduce-scatter of the partial gradients by P2P loads, projection backward
    + Adam on the owned 1/world slice (sharded optimiser state), all-gather of the updated parameters
    by P2P stores.  The fitter's xyz / cov / rgb / grads are re-homed in symmetric memory
    (torch.distributed._symmetric_memory: same allocation on every rank, peer pointers exchanged once);
    the only per-step synchronisation is the signal-pad barrier before and after the kernel.
    The squared-error partials are NOT exchanged per step: `stats()` all-reduces them when asked.
    """

    def __init__(self, fit, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib

        self.fit, self.group = fit, group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise ValueError("peer-memory exchange is for the GPUs of one NVSwitch box (<= 8)")
        name = self.group.group_name
        symm_mem.enable_symm_mem_for_group(name)
        fit.sync_params()
        n = fit.cur_num_points
        src = {"xyz": fit._t_xyz, "cov": fit._t_cov2d, "rgb": fit._t_f_dc, "grads": fit.grads}
        self.tensors, self.handles, self.tables = {}, {}, {}
        for key, t in src.items():
            s = symm_mem.empty(*t.shape, dtype=torch.float32, device=fit.device)
            s.copy_(t)
            self.tensors[key] = s
            self.handles[key] = symm_mem.rendezvous(s, name)
            ptrs = [int(p) for p in self.handles[key].buffer_ptrs]
            assert ptrs[self.rank] == s.data_ptr()
            self.tables[key] = (C.c_void_p * self.world)(*ptrs)
        fit._t_xyz, fit._t_cov2d, fit._t_f_dc = self.tensors["xyz"], self.tensors["cov"], self.tensors["rgb"]
        fit.grads = self.tensors["grads"]
        fit.external_optimizer = True
        fit.params.external_optimizer = 1
        fit.use_graph = False
        fit._invalidate_graphs()
        fit._bind()
        fit.grad_hook = self
        self._lib, self._C = _lib, C
        dist.barrier(self.group)
        torch.cuda.synchronize(fit.device)

    def __call__(self, fit):
        import torch

        h = self.handles["grads"]
        h.barrier(channel=0)        # every peer has finished the backward of this step
        st = self._C.c_void_p(torch.cuda.current_stream(fit.device).cuda_stream)
        self._lib.check(fit.lib.gi2d_fit_exchange_adam(
            self._C.byref(fit.params), self._C.byref(fit.buffers), self.rank, self.world, self.tables["grads"],
            self.tables["xyz"], self.tables["cov"], self.tables["rgb"], st), "fit_exchange_adam")
        h.barrier(channel=1)        # every peer's stores into my parameters have landed

    def global_stats(self):
        """fit.stats() with the squared error summed over the bands (one small all-reduce, on demand)."""
        import math

        import torch
        import torch.distributed as dist

        from .fit import STAT_SSE, STAT_SSE_SLOTS

        st = self.fit.stats()
        sse = torch.tensor([st["sse"]], dtype=torch.float64, device=self.fit.device)
        dist.all_reduce(sse, group=self.group)
        st["sse"] = float(sse.item())
        st["mse"] = st["sse"] / (3.0 * self.fit.H * self.fit.W)
        st["psnr"] = 10 * math.log10(1.0 / st["mse"]) if st["mse"] > 0 else float("inf")
        return st
