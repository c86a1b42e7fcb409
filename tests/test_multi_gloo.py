"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank searches its shard with the oracle (allowed in tests),
the per-shard top-k are all-gathered through annb200.distributed and merged; the result must equal the unsharded
oracle answer.  The device merge kernel itself is covered by tests/test_gpu_ivf.py (sharded test)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from annb200 import distributed as D
from annb200 import datagen
from oracle import oracle as o


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _merge_reference(g_ids, g_dist, k):
    """(distance, id) merge of [parts, nq, k] padded results -- numpy statement of annb_merge_topk_dev."""
    parts, nq, _ = g_ids.shape
    out_i = np.full((nq, k), -1, np.int64)
    out_d = np.full((nq, k), np.inf, np.float32)
    for q in range(nq):
        cand = [(float(g_dist[p, q, j]), int(g_ids[p, q, j])) for p in range(parts) for j in range(k) if g_ids[p, q, j] >= 0]
        cand.sort()
        for j, (d, i) in enumerate(cand[:k]):
            out_i[q, j], out_d[q, j] = i, np.float32(d)
    return out_i, out_d


def _merge_shards_reference(g_ids, g_dist, k):
    """(distance, shard, position) merge -- numpy statement of annb_merge_shards_dev: a stable sort on distance alone of the
    shards' lists concatenated in shard order."""
    parts, nq, _ = g_ids.shape
    out_i = np.full((nq, k), -1, np.int64)
    out_d = np.full((nq, k), np.inf, np.float32)
    for q in range(nq):
        ids = g_ids[:, q, :].reshape(-1)
        d = g_dist[:, q, :].reshape(-1)
        keep = ids >= 0
        order = np.argsort(d[keep], kind="stable")[:k]
        out_i[q, :order.size], out_d[q, :order.size] = ids[keep][order], d[keep][order]
    return out_i, out_d


def _worker(rank, world, port, kind, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = datagen.gaussian_noise(3000, 16, seed=77)
        q = datagen.subsample_with_noise(data, 40, seed=77)
        k = 7
        if kind == "flat":
            lo, hi = D.row_range(data.shape[0], world, rank)
            ix = o.build_flat(data[lo:hi], o.L2)
            ids, d, _ = o.flat_search(ix, q, k)
            ids = np.where(ids >= 0, ids + lo, ids)            # id_base of the shard
        else:
            full = o.build_ivf(data, o.L2, nlist=24, kmeans_iters=4)
            lb, le = D.list_ranges(full.offsets, world)[rank]
            nprobe = 5
            ids = np.full((q.shape[0], k), -1, np.int64)
            d = np.full((q.shape[0], k), np.inf, np.float32)
            for qi in range(q.shape[0]):
                cd = np.array([o.euclid_f32(q[qi], c) for c in full.centroids], np.float32)
                probed = o.select_probed(cd, np.arange(full.nlist), full.offsets, nprobe, k)     # global list sizes
                cand = []
                for c in probed:
                    if lb <= c < le:                                                              # only lists this rank owns
                        for v in range(int(full.offsets[c]), int(full.offsets[c + 1])):
                            cand.append((np.float32(o.euclid_f32(full.vectors[v], q[qi])), v))
                cand.sort()
                for j, (dd, v) in enumerate(cand[:k]):
                    ids[qi, j], d[qi, j] = full.original_ids[v], dd
        g_ids, g_d = D.allgather_topk(torch.from_numpy(ids), torch.from_numpy(d))
        m_ids, m_d = _merge_reference(g_ids.numpy(), g_d.numpy(), k)
        # the one-collective form used by ShardedSearch: interleaved [ids | distances] blocks, merged in shard order
        gathered = D.allgather_blocks(D.pack_block(torch.from_numpy(ids), torch.from_numpy(d)))
        b_ids, b_d = D.unpack_blocks(gathered, world, q.shape[0], k)
        assert torch.equal(b_ids, g_ids) and torch.equal(b_d.view(torch.int32), g_d.view(torch.int32))
        s_ids, s_d = _merge_shards_reference(b_ids.numpy(), b_d.numpy(), k)
        if rank == 0:
            ret["ids"], ret["dist"] = m_ids, m_d
            ret["s_ids"], ret["s_dist"] = s_ids, s_d
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["flat", "ivf"])
def test_two_rank_sharded_search_equals_unsharded(kind):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), kind, ret), nprocs=world, join=True)
    data = datagen.gaussian_noise(3000, 16, seed=77)
    q = datagen.subsample_with_noise(data, 40, seed=77)
    if kind == "flat":
        ref = o.flat_search(o.build_flat(data, o.L2), q, 7)
    else:
        ref = o.ivf_search(o.build_ivf(data, o.L2, nlist=24, kmeans_iters=4), q, 7, nprobe=5)
    assert (np.asarray(ret["dist"]).view(np.uint32) == ref[1].view(np.uint32)).all()
    assert (np.sort(np.asarray(ret["ids"]), axis=1) == np.sort(ref[0], axis=1)).all()
    # the shard-order merge reproduces the unsharded rows exactly, ids in the reference's own tie order included
    assert (np.asarray(ret["s_dist"]).view(np.uint32) == ref[1].view(np.uint32)).all()
    assert (np.asarray(ret["s_ids"]) == ref[0]).all()


def test_partition_helpers():
    assert [D.row_range(10, 3, r) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]
    off = np.array([0, 5, 5, 40, 41, 80, 100])
    for world in (1, 2, 4, 8):
        rs = D.list_ranges(off, world)
        assert rs[0][0] == 0 and rs[-1][1] == 6 and all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
    a, b = D.list_ranges(off, 2)
    # a boundary can be off by at most one list: imbalance <= 2 * the largest list
    assert abs((off[a[1]] - off[a[0]]) - (off[b[1]] - off[b[0]])) <= 2 * np.diff(off).max()
