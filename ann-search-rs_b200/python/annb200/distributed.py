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
