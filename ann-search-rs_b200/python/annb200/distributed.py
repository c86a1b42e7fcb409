"""Multi-GPU plumbing: one process per GPU (torch.distributed), database rows (flat) or inverted lists (IVF)
sharded across ranks, per-shard top-k all-gathered and merged under the (distance, id) order.

The exchange is the only collective of the path: `8 * nq * k` bytes per rank.  The search and the
merge run in libannb200 (`annb_*_search_dev`, `annb_merge_topk_dev`); torch supplies device buffers, the NCCL
all-gather and streams only.  The partition helpers are pure functions so they can be tested without a GPU.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def row_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Flat sharding: rank r owns rows [r*n/world, (r+1)*n/world)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def list_ranges(offsets, world: int) -> List[Tuple[int, int]]:
    """IVF sharding: contiguous list ranges with (nearly) equal vector counts, so every shard's slab stays one
    contiguous array and the scan work is balanced.  Returns [(list_begin, list_end)] per rank; ranges cover
    [0, nlist) without gaps and may be empty when there are more ranks than lists."""
    off = np.asarray(offsets, dtype=np.int64)
    nlist = off.size - 1
    n = int(off[-1])
    bounds = [0]
    for r in range(1, world):
        b = int(np.searchsorted(off, (n * r) / world, side="left"))
        bounds.append(min(max(b, bounds[-1]), nlist))
    bounds.append(nlist)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def allgather_topk(ids, dist, group=None):
    """All-gather per-shard results.  ids [nq, k] int64, dist [nq, k] float32 (torch tensors on the rank's device,
    or CPU tensors with the gloo backend) -> ([world, nq, k], [world, nq, k])."""
    import torch
    import torch.distributed as dist_
    world = dist_.get_world_size(group)
    g_ids = torch.empty((world,) + tuple(ids.shape), dtype=ids.dtype, device=ids.device)
    g_dist = torch.empty((world,) + tuple(dist.shape), dtype=dist.dtype, device=dist.device)
    dist_.all_gather_into_tensor(g_ids.view(-1), ids.contiguous().view(-1), group=group)
    dist_.all_gather_into_tensor(g_dist.view(-1), dist.contiguous().view(-1), group=group)
    return g_ids, g_dist


def merge_topk_device(g_ids, g_dist, out_ids=None, out_dist=None, stream=None):
    """annb_merge_topk_dev on CUDA tensors [parts, nq, k] -> ([nq, k], [nq, k])."""
    import torch

    from . import _check, lib
    parts, nq, k = g_ids.shape
    if out_ids is None:
        out_ids = torch.empty((nq, k), dtype=g_ids.dtype, device=g_ids.device)
    if out_dist is None:
        out_dist = torch.empty((nq, k), dtype=g_dist.dtype, device=g_dist.device)
    st = (stream or torch.cuda.current_stream(g_ids.device)).cuda_stream
    _check(lib().annb_merge_topk_dev(g_ids.data_ptr(), g_dist.data_ptr(), parts, nq, k, out_ids.data_ptr(), out_dist.data_ptr(), None, st))
    return out_ids, out_dist


def probe_pitch(nprobe: int) -> int:
    """Probe-list pitch used for the exchange: room for the probe expansion of select_probed_clusters."""
    return ((max(1, nprobe) + 32 + 31) // 32) * 32


def shard_block_bytes(nq: int, k: int) -> int:
    """Size of one shard's interleaved result block: [nq * k] int64 ids, [nq * k] float32 distances, [nq] float32 bounds of
    the shard-mode certificate and one int32 status word, padded to 256 bytes so every rank's slot of the gathered buffer
    stays aligned."""
    return ((nq * k * 12 + nq * 4 + 4 + 255) // 256) * 256


def pack_block(ids, dist):
    """One shard's interleaved result block (uint8 tensor of shard_block_bytes): ids [nq, k] int64, dist [nq, k] float32."""
    import torch
    nq, k = ids.shape
    blk = torch.zeros((shard_block_bytes(nq, k),), dtype=torch.uint8, device=ids.device)
    blk[:nq * k * 8].view(torch.int64).copy_(ids.contiguous().view(-1))
    blk[nq * k * 8:nq * k * 12].view(torch.float32).copy_(dist.contiguous().view(-1))
    return blk


def unpack_blocks(gathered, world: int, nq: int, k: int):
    """[world * block] uint8 -> (ids [world, nq, k] int64, dist [world, nq, k] float32) views-by-copy."""
    import torch
    b = shard_block_bytes(nq, k)
    g = gathered.view(world, b)
    ids = torch.stack([g[p, :nq * k * 8].view(torch.int64).view(nq, k) for p in range(world)])
    dist = torch.stack([g[p, nq * k * 8:nq * k * 12].view(torch.float32).view(nq, k) for p in range(world)])
    return ids, dist


def allgather_blocks(block, group=None):
    """ONE collective for ids and distances: all-gather of the ranks' interleaved blocks (any backend)."""
    import torch
    import torch.distributed as dist_
    world = dist_.get_world_size(group)
    out = torch.empty((world * block.numel(),), dtype=torch.uint8, device=block.device)
    dist_.all_gather_into_tensor(out, block, group=group)
    return out


class ShardedSearch:
    """One sharded search step per call, one process per GPU.  `index` is this rank's shard -- a row range
    (flat, built with id_base = first row) or a list range (IVF) -- and ranks hold ascending ranges in rank order.

        flat : shard-mode search of the whole batch (annb_flat_search_shard_dev) -> ONE all-gather of the interleaved
               [ids | distances] blocks (12 * nq * k bytes per rank) -> annb_merge_shards_dev.
        IVF  : every rank ranks the centroids for ITS slice of the batch (annb_ivf_route_dev), ONE all-gather of the
               slices' [probe lists | probe counts | status] blocks, every rank scans its own lists for the whole batch
               (annb_ivf_search_probes_shard_dev), then the same result exchange.
        certificate : shards do not certify their own k-th neighbour -- it is usually irrelevant to the merged top-k --
               but report a bound; after the merge every rank tests its bound against the merged k-th distances
               (annb_shard_check_dev), one 4-byte all-reduce tells all ranks whether anybody has to refine
               (annb_shard_refine_dev: exact kernels for the listed queries), and only then the exchange is repeated.

    Everything is enqueued on `stream` (default: torch's current stream), collectives included.  A rank whose library
    call fails still enters every collective of the step; the gathered status words carry the error, and all ranks raise
    together instead of leaving the others blocked in NCCL.

    defer=True (a serving loop): the call returns as soon as the step is enqueued -- the verdict words travel to pinned host
    memory behind it -- and `resolve()` finishes the step later: it waits for the verdict and, if some shard has to refine,
    runs the refine + second exchange (a collective path: every rank reaches the same verdict from the gathered bounds, so all
    ranks take it together).  The returned tensors are final only after resolve(); an object holds one pending step, so a
    loop that wants the host one step ahead of the devices alternates two objects (bench.py does)."""

    def __init__(self, index, nq: int, dim: int, k: int, nprobe: int = 0, group=None, device=None):
        import torch
        import torch.distributed as dist_
        self.index, self.nq, self.dim, self.k, self.nprobe, self.group = index, nq, dim, k, nprobe, group
        self.world, self.rank = dist_.get_world_size(group), dist_.get_rank(group)
        dev = device or torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.block = shard_block_bytes(nq, k)
        self.mine = torch.empty((self.block,), dtype=torch.uint8, device=dev)
        self.gathered = torch.empty((self.world * self.block,), dtype=torch.uint8, device=dev)
        self.ids = self.mine[:nq * k * 8].view(torch.int64).view(nq, k)
        self.dist = self.mine[nq * k * 8:nq * k * 12].view(torch.float32).view(nq, k)
        self.out_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        self.out_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
        self.bound = self.mine[nq * k * 12:nq * k * 12 + nq * 4].view(torch.float32)          # travels with the block
        self.status = self.mine[nq * k * 12 + nq * 4:nq * k * 12 + nq * 4 + 4].view(torch.int32)   # 1 = this rank's library call failed
        self.status.zero_()
        self.refined_queries = 0          # cumulative: queries this rank recomputed exactly after the merged check
        self.h_verdict = torch.zeros((2,), dtype=torch.int32).pin_memory()   # deferred steps: {mine, any} land here
        self.ev = torch.cuda.Event()
        self._pending = None              # (queries, stream object, error) of a deferred step awaiting resolve()
        self.is_ivf = bool(index.info().is_ivf)
        if self.is_ivf:
            self.pitch = probe_pitch(nprobe or int(max(1, index.info().nlist ** 0.5)))
            self.per = (nq + self.world - 1) // self.world
            # one exchange block per rank: [per * pitch] probe cells, [per] probe counts, [1] status word
            self.route_words = self.per * self.pitch + self.per + 1
            self.my_route = torch.zeros((self.route_words,), dtype=torch.int32, device=dev)
            self.all_route = torch.empty((self.world * self.route_words,), dtype=torch.int32, device=dev)
            self.probes = torch.empty((self.world * self.per, self.pitch), dtype=torch.int32, device=dev)
            self.nprobes = torch.empty((self.world * self.per,), dtype=torch.int32, device=dev)

    def _exchange_and_merge(self, L, st):
        import torch.distributed as dist_
        dist_.all_gather_into_tensor(self.gathered, self.mine, group=self.group)
        return L.annb_merge_shards_dev(self.gathered.data_ptr(), self.block, self.nq * self.k * 8, self.world, self.nq, self.k, self.out_ids.data_ptr(),
                                       self.out_dist.data_ptr(), None, st)

    def __call__(self, queries, stream=None, defer=False):
        """queries: [nq, dim] float32 CUDA tensor (the whole batch, on every rank).  Returns (ids, dist) on the device
        (defer=True: final only after resolve(); `queries` must stay untouched until then)."""
        import ctypes as C

        import torch
        import torch.distributed as dist_

        from . import AnnSearchError, lib
        L = lib()
        self.resolve()
        st_obj = stream or torch.cuda.current_stream(self.dev)
        st = st_obj.cuda_stream
        nq, dim, k = self.nq, self.dim, self.k
        err = None

        def failed(rc):
            return AnnSearchError(rc, L.annb_last_error().decode("utf-8", "replace"))

        with torch.cuda.stream(st_obj):
            if self.is_ivf:
                lo, hi = min(nq, self.rank * self.per), min(nq, (self.rank + 1) * self.per)
                pw = self.per * self.pitch
                if hi > lo:
                    rc = L.annb_ivf_route_dev(self.index.handle, queries[lo:hi].data_ptr(), hi - lo, dim, k, self.nprobe, self.my_route.data_ptr(),
                                              self.my_route[pw:].data_ptr(), self.pitch, st)
                    if rc != 0:   # e.g. Unsupported: a probe set did not fit the pitch -- keep the collectives matched, raise at the verdict
                        err = failed(rc)
                dist_.all_gather_into_tensor(self.all_route, self.my_route, group=self.group)
                blocks = self.all_route.view(self.world, self.route_words)
                self.probes.view(self.world, pw).copy_(blocks[:, :pw])
                self.nprobes.view(self.world, self.per).copy_(blocks[:, pw:pw + self.per])
                if err is None:
                    rc = L.annb_ivf_search_probes_shard_dev(self.index.handle, queries.data_ptr(), nq, dim, k, self.nprobe, self.probes.data_ptr(),
                                                            self.nprobes.data_ptr(), self.pitch, self.ids.data_ptr(), self.dist.data_ptr(),
                                                            self.bound.data_ptr(), st)
                    err = failed(rc) if rc != 0 else None
            else:
                rc = L.annb_flat_search_shard_dev(self.index.handle, queries.data_ptr(), nq, dim, k, self.ids.data_ptr(), self.dist.data_ptr(),
                                                  self.bound.data_ptr(), st)
                err = failed(rc) if rc != 0 else None
            self.status.fill_(1 if err is not None else 0)
            # Every rank holds every shard's bounds and status after the exchange: the verdict needs no further collective.
            if defer:      # merge and verdict in one pass over the gathered blocks, verdict words on their way to pinned memory
                dist_.all_gather_into_tensor(self.gathered, self.mine, group=self.group)
                rc = L.annb_merge_check_shards_async_dev(self.index.handle, self.gathered.data_ptr(), self.block, nq * k * 8, nq * k * 12, self.world,
                                                         self.rank, nq, k, self.out_ids.data_ptr(), self.out_dist.data_ptr(), self.h_verdict.data_ptr(), st)
                if rc != 0 and err is None:
                    err = failed(rc)
                self.ev.record(st_obj)
                self._pending = (queries, st_obj, err)
                return self.out_ids, self.out_dist
            rc = self._exchange_and_merge(L, st)
            if rc != 0 and err is None:
                err = failed(rc)
            self._finish(queries, st_obj, err)
        return self.out_ids, self.out_dist

    def resolve(self):
        """Finish a deferred step (no-op when none is pending).  Collective when a refine is needed: call it on all ranks."""
        if self._pending is None:
            return
        import torch
        queries, st_obj, err = self._pending
        self._pending = None
        self.ev.synchronize()
        mine, any_ = int(self.h_verdict[0]), int(self.h_verdict[1])
        if err is not None or (any_ & 2):
            raise err if err is not None else RuntimeError("another rank failed in the sharded search step")
        if any_ & 1:
            # the handle may have searched since: the synchronous check rebuilds the list of queries to refine, then the usual path
            with torch.cuda.stream(st_obj):
                self._finish(queries, st_obj, None)

    def _finish(self, queries, st_obj, err):
        """Synchronous verdict of the step whose blocks are in self.gathered / self.out_dist, refine + second exchange if needed."""
        import ctypes as C

        from . import AnnSearchError, lib
        L = lib()
        st = st_obj.cuda_stream
        nq, dim, k = self.nq, self.dim, self.k

        def failed(rc):
            return AnnSearchError(rc, L.annb_last_error().decode("utf-8", "replace"))

        mine, any_ = C.c_uint32(0), C.c_uint32(0)
        rc = L.annb_shard_check_gathered_dev(self.index.handle, self.gathered.data_ptr(), self.block, nq * k * 12, self.world, self.rank,
                                             self.out_dist.data_ptr(), nq, k, C.byref(mine), C.byref(any_), st)      # (one read-back: synchronises)
        if rc != 0 and err is None:
            err = failed(rc)
        if err is not None or (any_.value & 2):
            raise err if err is not None else RuntimeError("another rank failed in the sharded search step")
        if any_.value & 1:
            if mine.value:
                rc = L.annb_shard_refine_dev(self.index.handle, queries.data_ptr(), nq, dim, k, self.nprobe,
                                             self.probes.data_ptr() if self.is_ivf else None, self.nprobes.data_ptr() if self.is_ivf else None,
                                             self.pitch if self.is_ivf else 0, self.ids.data_ptr(), self.dist.data_ptr(), st)
                if rc != 0:
                    err = failed(rc)
                self.refined_queries += int(mine.value)
            self.status.fill_(1 if err is not None else 0)
            rc = self._exchange_and_merge(L, st)
            if rc == 0:     # (status words only: the refined rows are exact)
                rc = L.annb_shard_check_gathered_dev(self.index.handle, self.gathered.data_ptr(), self.block, nq * k * 12, self.world, self.rank,
                                                     self.out_dist.data_ptr(), nq, k, C.byref(mine), C.byref(any_), st)
            if err is not None or rc != 0 or (any_.value & 2):
                raise err if err is not None else RuntimeError("a rank failed while refining the sharded search step")


def ivf_search_sharded(index, queries, k: int, nprobe: int, out_ids=None, out_dist=None, group=None, stream=None):
    """One sharded IVF search step (convenience wrapper over ShardedSearch; allocates its buffers per call).
    Returns (ids [nq, k] int64, dist [nq, k] float32) on the rank's device; raises on ALL ranks if any rank failed."""
    nq, dim = queries.shape
    step = ShardedSearch(index, nq, dim, k, nprobe, group, queries.device)
    ids, dst = step(queries, stream)
    if out_ids is not None:
        out_ids.copy_(ids)
        ids = out_ids
    if out_dist is not None:
        out_dist.copy_(dst)
        dst = out_dist
    return ids, dst
