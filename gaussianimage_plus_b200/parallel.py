"""Multi-GPU partitioning of the hot path.  One process per GPU, torch.distributed for
the plumbing; the reference itself is single-GPU (`train.py:39`), so both modes are new design.

1. Image sets (Kodak-24, DIV2K-100): images are independent units -> `shard_images` deals image i to
   rank i mod world; NO collective on the data path; `gather_metrics` collects the per-image scalars
   once at the end (the averages of train.py:327-340).

2. One very large image: `TileRowPartition` gives rank r a contiguous band of tile rows.
   * `TileRowFit` (the product path): sharded projection + sharded optimiser, the exchange fused into the
     step's last kernel over NVLink peer memory with in-kernel flag synchronisation (gi2d_tilerow_step).
   * `TileRowPartition.make_grad_hook()` (the library baseline it is measured against): parameters and
     Adam state replicated, every rank projects every Gaussian, ONE NCCL all-reduce(SUM) of the packed
     f32[N,8] gradient buffer (+ the 64 squared-error partials) per iteration on the kernels' stream.
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


class TileRowFit:
    """ONE image fitted by `world` ranks, split by tile rows (BASELINE.json configs[4]; gi2d_tilerow_step).

    Rank q rasterizes the band `partition.band(q)`; every Gaussian has one OWNER rank (equal contiguous
    slices) that alone keeps its parameters and Adam moments, applies the optimiser and projects it.  Per
    step the owner pulls the partial gradient rows of the ranks whose band the Gaussian's tile box overlaps
    (P2P loads) and pushes the new projected record + box to the ranks that need it (P2P stores); cross-GPU
    ordering is a pair of flag words per peer inside the kernels, so a step is ONE C call (4 + 1 kernels), has
    no host synchronisation and replays from a CUDA graph.  Peer-visible buffers (gradients, records, boxes,
    flags) live in ONE symmetric-memory allocation per rank (torch.distributed._symmetric_memory: identical
    allocation on every rank, peer pointers exchanged once at set-up).

    `fit` is this rank's GaussianImageFitter constructed with tile_rows=partition.band(rank) and the SAME
    initial parameters on every rank.  After the run `gather_params()` collects the owned slices.

    `TileRowFit.emulate(fits, partition)` builds the same thing for `world` fitters on ONE GPU (ordinary
    device memory, in-kernel flags off, the ranks stepped one after the other on one stream): the data path
    -- band clipping, reduce by box overlap, sharded Adam, scatter of records -- is then testable on a
    single-GPU box.
    """

    def __init__(self, fit, partition: "TileRowPartition", group=None, _emulated=None):
        import ctypes as C

        import torch

        from . import _lib

        self.fit, self.partition = fit, partition
        self._lib, self._C = _lib, C
        n = fit.cur_num_points
        if fit.loss_w[2] != 0 or fit.loss_ms[0] != 0:
            raise ValueError("SSIM losses are not available for a tile-row band (the window crosses band borders)")
        if _emulated is None:
            import torch.distributed as dist

            self.group = group if group is not None else dist.group.WORLD
            self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        else:
            self.group = None
            self.rank, self.world = _emulated
        if self.world > _lib.MAX_RANKS:
            raise ValueError("the peer-memory exchange is for the GPUs of one NVSwitch box (<= 8)")
        if tuple(fit.tile_rows) != tuple(partition.band(self.rank)):
            raise ValueError("construct the fitter with tile_rows=partition.band(rank)")
        # ONE peer-visible allocation: grads f32[N,8] | proj f32[N,8] | boxes u16[N,4] | flags u32[16]
        words = n * 8 + n * 8 + n * 2 + 64
        if _emulated is None:
            import warnings

            import torch.distributed._symmetric_memory as symm_mem

            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                try:
                    symm_mem.enable_symm_mem_for_group(self.group.group_name)
                except Exception:
                    pass
                self.shared = symm_mem.empty(words, dtype=torch.float32, device=fit.device)
                self.shared.zero_()
                self._handle = symm_mem.rendezvous(self.shared, self.group.group_name)
            bases = [int(ptr) for ptr in self._handle.buffer_ptrs]
            assert bases[self.rank] == self.shared.data_ptr()
        else:
            self.shared = torch.zeros(words, dtype=torch.float32, device=fit.device)
            bases = None   # filled in by emulate()
        self.ctrl = torch.zeros(8, dtype=torch.int32, device=fit.device)
        fit.sync_params()
        fit.grads = self.shared[0:n * 8].view(n, 8)
        fit.proj = self.shared[n * 8:n * 16].view(n, 8)
        fit._keep_exchange_buffers = True      # _alloc_state() must not move them (ensure_capacity regrows)
        fit.external_optimizer = 2
        fit.params.external_optimizer = 2
        fit.track_best = False                 # the parameters are sharded: no per-rank best-state snapshot
        fit.grad_hook = None
        fit._invalidate_graphs()
        fit._bind()
        per = (n + self.world - 1) // self.world
        self.own = (min(n, self.rank * per), min(n, (self.rank + 1) * per))
        self._graphs = {}
        self._bases = bases
        if bases is not None:
            self._finish_setup()

    def _finish_setup(self):
        import torch

        C, _lib, fit, n = self._C, self._lib, self.fit, self.fit.cur_num_points
        tr = _lib.TileRow()
        tr.rank, tr.world = self.rank, self.world
        for q in range(self.world + 1):
            tr.band_edge[q] = self.partition.edges[q]
        tr.own_begin, tr.own_end = self.own
        tr.sync = 1 if self.group is not None else 0
        for q, base in enumerate(self._bases):
            tr.peer_grads[q] = base
            tr.peer_proj[q] = base + n * 32
            tr.peer_boxes[q] = base + n * 64
            tr.peer_flags[q] = base + n * 72
        tr.ctrl = self.ctrl.data_ptr()
        self.tr = tr
        torch.cuda.synchronize(fit.device)     # (the zeroing of self.shared has completed)
        if self.group is not None:
            import torch.distributed as dist

            dist.barrier(self.group)           # every rank's buffers exist and are zeroed
        with torch.cuda.device(fit.device):
            _lib.check(fit.lib.gi2d_tilerow_init(C.byref(fit.params), C.byref(fit.buffers), C.byref(tr),
                                                 self._stream()), "tilerow_init")
        torch.cuda.synchronize(fit.device)
        if self.group is not None:
            import torch.distributed as dist

            dist.barrier(self.group)           # every rank's records have landed everywhere

    @classmethod
    def emulate(cls, fits, partition):
        """`len(fits)` ranks on ONE GPU (tests): ordinary device memory, kernels ordered by the stream."""
        world = len(fits)
        objs = [cls(f, partition, _emulated=(r, world)) for r, f in enumerate(fits)]
        bases = [o.shared.data_ptr() for o in objs]
        import torch

        torch.cuda.synchronize(fits[0].device)
        for o in objs:
            o._bases = bases
        # init in two passes: every rank's zeroing must precede any rank's scatter
        for o in objs:
            o.shared.zero_()
        for o in objs:
            o._finish_setup()
        for o in objs:
            o._peers = objs
        return objs

    def _stream(self):
        import torch

        return self._C.c_void_p(torch.cuda.current_stream(self.fit.device).cuda_stream)

    def _enqueue(self, phase: int, with_backward: int = 1):
        fit = self.fit
        self._lib.check(fit.lib.gi2d_tilerow_step(self._C.byref(fit.params), self._C.byref(fit.buffers),
                                                  self._C.byref(self.tr), with_backward, phase, self._stream()),
                        "tilerow_step")

    def train_iter(self):
        """One iteration, asynchronous.  (Emulated ranks: call `TileRowFit.step_all(objs)` instead.)"""
        if self.group is None:
            raise RuntimeError("emulated ranks are stepped together: TileRowFit.step_all(objs)")
        import torch

        with torch.cuda.device(self.fit.device):
            self._enqueue(3)
        self.fit._expected_step += 1

    def train_iters(self, n: int, unroll: int = 4):
        """`n` iterations; replayed from a CUDA graph of `unroll` steps when the fitter allows graphs (the epoch
        and the flags live in device memory, so a replay needs no per-step host input)."""
        import torch

        fit = self.fit
        if self.group is None:
            raise RuntimeError("emulated ranks are stepped together: TileRowFit.step_all(objs)")
        with torch.cuda.device(fit.device):
            if not fit.use_graph or n < unroll:
                for _ in range(n):
                    self._enqueue(3)
            else:
                key = (fit.gt_hwc.data_ptr(), fit.isect_capacity, unroll)
                g = self._graphs.get(key)
                if g is None:
                    for _ in range(2):          # load the kernels un-captured
                        self._enqueue(3)
                    n -= 2
                    g = torch.cuda.CUDAGraph()
                    s = torch.cuda.Stream(device=fit.device)
                    s.wait_stream(torch.cuda.current_stream(fit.device))
                    with torch.cuda.stream(s):
                        with torch.cuda.graph(g, stream=s):
                            for _ in range(unroll):
                                self._enqueue(3)
                    torch.cuda.current_stream(fit.device).wait_stream(s)
                    self._graphs = {key: g}
                for _ in range(max(n, 0) // unroll):
                    g.replay()
                for _ in range(max(n, 0) % unroll):
                    self._enqueue(3)
        fit._expected_step += n if n > 0 else 0

    @staticmethod
    def step_all(objs, n: int = 1):
        """Emulated ranks on one GPU: the band steps of every rank, then the exchanges of every rank."""
        for _ in range(n):
            for o in objs:
                o._enqueue(1)
            for o in objs:
                o._enqueue(2)
            for o in objs:
                o.fit._expected_step += 1

    def check(self):
        """Raises when a cross-GPU flag wait timed out inside a kernel (a peer died or never launched)."""
        if int(self.ctrl[3].item()) != 0:
            raise self._lib.Gi2dError("tile-row exchange: a flag wait timed out (peer not running?)")

    def local_stats(self) -> dict:
        return self.fit.stats()

    def stats(self) -> dict:
        """fit.stats() with the squared error and the intersections summed over the bands (one small all-reduce
        on demand; emulated ranks: summed over the peer objects).  step / lr / overflow are identical on all
        ranks by construction."""
        import math

        import torch

        st = self.fit.stats()
        part = torch.tensor([st["sse"], float(st["num_intersects"])], dtype=torch.float64, device=self.fit.device)
        if self.group is not None:
            import torch.distributed as dist

            dist.all_reduce(part, group=self.group)
        else:
            part = sum(torch.tensor([o.fit.stats()["sse"], float(o.fit.stats()["num_intersects"])],
                                    dtype=torch.float64, device=self.fit.device) for o in self._peers)
        st["band_num_intersects"] = st["num_intersects"]
        st["sse"], st["num_intersects"] = float(part[0].item()), int(part[1].item())
        st["mse"] = st["sse"] / (3.0 * self.fit.H * self.fit.W)
        st["loss"] = st["mse"]
        st["psnr"] = 10 * math.log10(1.0 / st["mse"]) if st["mse"] > 0 else float("inf")
        return st

    def catch_up(self) -> dict:
        """Re-run the iterations that were no-ops because SOME rank's band overflowed its intersection buffers
        (the veto is global, so every rank is missing the same number of steps): the overflowing rank regrows,
        everybody re-runs.  Collective; synchronises."""
        st = self.fit.stats()
        while st["step"] < self.fit._expected_step:
            lost = self.fit._expected_step - st["step"]
            # this rank's band overflowed: the whole buffer (scan + placement) or one tile's bucket (max_tile is set
            # by the band step only when a bucket of THIS rank was too small)
            if st["num_intersects"] > self.fit.isect_capacity or st.get("max_tile", 0) > 0:
                expected = self.fit._expected_step
                tiles = self.fit.tile_bounds[0] * self.fit.tile_bounds[1]
                self.fit._capacity_hint = max(int(st["num_intersects"] * (2 if st.get("max_tile", 0) else 4)),
                                              int(st.get("max_tile", 0) * 1.25 + 8) * tiles)
                self.fit._step0 = st["step"]
                self.fit._alloc_state(zero_moments=False)
                self.fit._bind()
                self.fit._expected_step = expected
                self._graphs = {}
            self.fit._expected_step -= lost
            if self.group is not None:
                self.train_iters(lost)
            else:
                return st   # emulated: the caller steps all ranks (step_all) after every rank regrew
            st = self.fit.stats()
        return st

    def gather_params(self):
        """(xyz, cov2d, features_dc) of the whole model on every rank: all-gather of the owned slices."""
        import torch

        fit = self.fit
        outs = []
        for t in (fit._t_xyz, fit._t_cov2d, fit._t_f_dc):
            full = t.clone()
            if self.group is not None:
                import torch.distributed as dist

                per = (fit.cur_num_points + self.world - 1) // self.world
                pad = torch.zeros(per * self.world, t.shape[1], dtype=t.dtype, device=t.device)
                mine = torch.zeros(per, t.shape[1], dtype=t.dtype, device=t.device)
                mine[:self.own[1] - self.own[0]] = t[self.own[0]:self.own[1]]
                dist.all_gather_into_tensor(pad, mine, group=self.group)
                full = pad[:fit.cur_num_points].clone()
            else:
                for o in self._peers:
                    src = {"xyz": o.fit._t_xyz, "cov": o.fit._t_cov2d, "rgb": o.fit._t_f_dc}
                    which = "xyz" if t is fit._t_xyz else ("cov" if t is fit._t_cov2d else "rgb")
                    full[o.own[0]:o.own[1]] = src[which][o.own[0]:o.own[1]]
            outs.append(full)
        return tuple(outs)

    def verify_records(self) -> int:
        """Collective check of the exchange (bench.py asserts it in-run): for every Gaussian whose tile box -- in
        the owner's own copy -- overlaps this rank's band, this rank's copy of the projected record and of the box
        must be bit-identical to the owner's.  Returns the number of mismatching rows summed over the ranks."""
        import torch
        import torch.distributed as dist

        fit, n = self.fit, self.fit.cur_num_points
        torch.cuda.synchronize(fit.device)
        my_boxes = self.shared[n * 16:n * 18].view(torch.int16).view(n, 4)
        per = (n + self.world - 1) // self.world
        lo, hi = self.partition.band(self.rank)
        bad = torch.zeros(1, dtype=torch.int64, device=fit.device)
        for q in range(self.world):
            g0, g1 = min(n, q * per), min(n, (q + 1) * per)
            rec = fit.proj[g0:g1].clone() if q == self.rank else torch.empty(g1 - g0, 8, device=fit.device)
            box = my_boxes[g0:g1].clone() if q == self.rank else torch.empty(g1 - g0, 4, dtype=torch.int16, device=fit.device)
            if self.group is not None:
                src = dist.get_global_rank(self.group, q)
                dist.broadcast(rec, src=src, group=self.group)
                dist.broadcast(box.view(torch.int32), src=src, group=self.group)   # (NCCL has no 16-bit integer type)
            b = box.to(torch.int32) & 0xFFFF
            hit = (b[:, 2] > b[:, 0]) & (b[:, 3] > lo) & (b[:, 1] < hi)
            same = (fit.proj[g0:g1].view(torch.int32) == rec.view(torch.int32)).all(dim=1) & \
                   (my_boxes[g0:g1] == box).all(dim=1)
            bad += (hit & ~same).sum()
        if self.group is not None:
            dist.all_reduce(bad, group=self.group)
        return int(bad.item())

    def nvlink_bytes_per_step(self) -> int:
        """Bytes this rank moves over NVLink in one step (counted from the boxes currently on the device): 32 per
        (owned Gaussian, remote rank whose band its box overlaps) loaded + 40 per such pair stored.  Synchronises."""
        import torch

        n = self.fit.cur_num_points
        boxes = self.shared[n * 16:n * 18].view(torch.int16).view(n, 4)[self.own[0]:self.own[1]].to(torch.int32) & 0xFFFF
        total = 0
        for q in range(self.world):
            if q == self.rank:
                continue
            lo, hi = self.partition.edges[q], self.partition.edges[q + 1]
            hit = (boxes[:, 2] > boxes[:, 0]) & (boxes[:, 3] > lo) & (boxes[:, 1] < hi)
            total += int(hit.sum().item()) * 72
        return total
