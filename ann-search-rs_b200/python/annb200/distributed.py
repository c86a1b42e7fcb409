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


def ivf_search_sharded(index, queries, k: int, nprobe: int, out_ids=None, out_dist=None, group=None, stream=None):
    """One sharded IVF search step on CUDA tensors (one process per GPU, `index` = this rank's list range):
      1. every rank ranks the centroids for ITS slice of the query batch (annb_ivf_route_dev),
      2. the probe lists are all-gathered (4 * nq * pitch bytes),
      3. every rank scans its own lists for the whole batch (annb_ivf_search_probes_dev),
      4. the per-shard top-k are all-gathered and merged (annb_merge_topk_dev).
    Raises AnnSearchError(Unsupported) when a probe set does not fit the pitch.
    Returns (ids [nq, k] int64, dist [nq, k] float32) on the rank's device."""
    import torch
    import torch.distributed as dist_

    from . import _check, lib
    world, rank = dist_.get_world_size(group), dist_.get_rank(group)
    nq, dim = queries.shape
    dev = queries.device
    st = (stream or torch.cuda.current_stream(dev)).cuda_stream
    L = lib()
    pitch = probe_pitch(nprobe)
    per = (nq + world - 1) // world
    lo, hi = min(nq, rank * per), min(nq, (rank + 1) * per)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    dst = torch.empty((nq, k), dtype=torch.float32, device=dev)
    my_probes = torch.full((per, pitch), -1, dtype=torch.int32, device=dev)
    my_n = torch.zeros((per,), dtype=torch.int32, device=dev)
    if hi > lo:
        # ANNB_ERR_UNSUPPORTED here means a probe set did not fit the pitch (tiny lists, huge k): the error is raised on this
        # rank before any collective of the step, so the job fails loudly instead of hanging; use a larger pitch then
        _check(L.annb_ivf_route_dev(index.handle, queries[lo:hi].data_ptr(), hi - lo, dim, k, nprobe, my_probes.data_ptr(), my_n.data_ptr(), pitch, st))
    g_probes = torch.empty((world * per, pitch), dtype=torch.int32, device=dev)
    g_n = torch.empty((world * per,), dtype=torch.int32, device=dev)
    dist_.all_gather_into_tensor(g_probes.view(-1), my_probes.view(-1), group=group)
    dist_.all_gather_into_tensor(g_n, my_n, group=group)
    _check(L.annb_ivf_search_probes_dev(index.handle, queries.data_ptr(), nq, dim, k, nprobe, g_probes.data_ptr(), g_n.data_ptr(), pitch,
                                        ids.data_ptr(), dst.data_ptr(), None, st))
    g_ids, g_dist = allgather_topk(ids, dst, group)
    return merge_topk_device(g_ids, g_dist, out_ids, out_dist, stream)
