#!/usr/bin/env python
"""One rank's share of a sharded search step, on ONE GPU, phase by phase (CUDA events).

The 8-GPU step is latency-bound; ncu cannot wrap a multi-rank command.  This tool builds the same shard a rank of a
`world`-way job holds (IVF: its list range; flat: its row range), runs the calls ShardedSearch makes -- route for the
rank's slice of the batch, shard-mode scan of the whole batch, merge of `world` result blocks, the merged check -- and
times each between CUDA events.  The collectives are not emulated (their payload is printed); under
`ncu --metrics gpu__time_duration.sum` the launch list of one such step is the per-rank kernel budget.

    python tools/shard_emulate.py --workload ivf --world 8 --nprobe 32
    python tools/shard_emulate.py --workload flat --world 8
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ann-search-rs_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ivf", choices=["ivf", "flat", "c5"])
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--nlist", type=int, default=4096)
    ap.add_argument("--nprobe", type=int, default=32)
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16", "sq8"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--option", action="append", default=[])
    args = ap.parse_args()
    import torch

    import annb200
    import gpu_setup as gs
    from annb200 import distributed as D
    L = annb200.lib()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    W, R = args.world, args.rank
    ivf = args.workload == "ivf"
    n = args.n or {"ivf": 10_000_000, "flat": 1_000_000, "c5": 2_000_000}[args.workload]
    dim = args.dim or (50 if args.workload == "c5" else 128)
    k = args.k or (15 if args.workload == "c5" else 10)
    nq = args.nq
    metric = annb200.COSINE if args.workload == "flat" else annb200.L2
    dt = {"f32": annb200.F32, "bf16": annb200.BF16, "sq8": annb200.SQ8}[args.dtype]
    data_t = gs.correlated_gpu(n, dim, dev, seed=42)
    q_t = data_t[:nq].contiguous() if args.workload == "c5" else gs.subsample_with_noise_gpu(data_t, nq, seed=42)
    if ivf:
        base = gs.build_ivf_parts_gpu(data_t, args.nlist, annb200.F32, 0, seed=42, kmeans_iters=8)
        del data_t
        parts = gs.requantise_parts(base, dt)
        lb, le = D.list_ranges(base["offsets"], W)[R]
        index = gs.ivf_handle_from_parts(parts, n, dim, dt, annb200.L2, 0, lb, le)
    else:
        r0, r1 = D.row_range(n, W, R)
        index = gs._flat_handle_from_device(data_t[r0:r1].contiguous(), metric, dt, 0, id_base=r0)
        del data_t
    torch.cuda.empty_cache()
    for kv in args.option:
        key, val = kv.split("=")
        index.set_option(key, int(val))
    st = torch.cuda.current_stream(dev).cuda_stream
    block = D.shard_block_bytes(nq, k)
    mine = torch.zeros((block,), dtype=torch.uint8, device=dev)
    gathered = torch.zeros((W * block,), dtype=torch.uint8, device=dev)
    ids = mine[:nq * k * 8].view(torch.int64).view(nq, k)
    dist = mine[nq * k * 8:nq * k * 12].view(torch.float32).view(nq, k)
    bound = mine[nq * k * 12:nq * k * 12 + nq * 4].view(torch.float32)
    out_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    out_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    per = (nq + W - 1) // W
    pitch = D.probe_pitch(args.nprobe)
    probes = torch.zeros((W * per, pitch), dtype=torch.int32, device=dev)
    nprobes = torch.zeros((W * per,), dtype=torch.int32, device=dev)
    if ivf:   # the other ranks' slices of the probe exchange (untimed)
        for r in range(W):
            lo, hi = min(nq, r * per), min(nq, (r + 1) * per)
            if hi > lo:
                annb200._check(L.annb_ivf_route_dev(index.handle, q_t[lo:hi].data_ptr(), hi - lo, dim, k, args.nprobe, probes[r * per:].data_ptr(),
                                                    nprobes[r * per:].data_ptr(), pitch, st))
    torch.cuda.synchronize()
    if ivf:
        # the scan's task list of this shard, replayed on the host: (list x 128-query group) tasks handed longest list first to
        # 148 persistent CTAs; cost model = tiles * t_tile + t_task.  Shows how much of the scan time is imbalance.
        import heapq
        off = np.asarray(base["offsets"], dtype=np.int64)
        pr = probes.cpu().numpy()[:nq]
        npr = nprobes.cpu().numpy()[:nq]
        mask = np.arange(pitch)[None, :] < npr[:, None]
        cells = pr[mask]
        cells = cells[(cells >= lb) & (cells < le)]
        cnt = np.bincount(cells, minlength=args.nlist)
        rows = (off[1:] - off[:-1])
        tiles = (rows + 127) // 128
        order = np.argsort(-rows[lb:le], kind="stable") + lb
        for tmax in (0, 32, 16, 8):
            tasks = []
            for c in order:
                g = (cnt[c] + 127) // 128
                if g == 0:
                    continue
                if tmax and tiles[c] > tmax:
                    parts_ = (tiles[c] + tmax - 1) // tmax
                    sub = [tiles[c] // parts_ + (1 if i < tiles[c] % parts_ else 0) for i in range(parts_)]
                else:
                    sub = [tiles[c]]
                for _ in range(g):
                    tasks.extend(sub)
            t_tile, t_task = 3.0, 6.0    # us (5 900 cycles per tile at 1.965 GHz; gather + drain per task)
            heap = [0.0] * 148
            heapq.heapify(heap)
            for t in tasks:                      # dynamic hand-out in list order = the atomic task counter
                heapq.heappush(heap, heapq.heappop(heap) + t * t_tile + t_task)
            total = sum(t * t_tile + t_task for t in tasks)
            print(f"[task replay] split at {tmax or 'none':>4} tiles: {len(tasks)} tasks, {int(sum(tasks))} tiles, longest task {max(tasks)} tiles, "
                  f"balanced {total / 148:.0f} us, makespan {max(heap):.0f} us; lists {le - lb}, rows per list min/median/max "
                  f"{rows[lb:le].min()}/{int(np.median(rows[lb:le]))}/{rows[lb:le].max()}, pairs per list median/max {int(np.median(cnt[lb:le]))}/{cnt[lb:le].max()}")
    names = ["route(slice)", "scan(shard)", "merge", "check"]
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 2)] for _ in range(args.steps)]
    h_verdict = torch.zeros((2,), dtype=torch.int32).pin_memory()
    tmp_p = torch.zeros((per * pitch,), dtype=torch.int32, device=dev)
    tmp_n = torch.zeros((per,), dtype=torch.int32, device=dev)
    mine_c, any_c = C.c_uint32(0), C.c_uint32(0)
    refine = 0
    for it in range(-args.warmup, args.steps):
        e = ev[max(it, 0)]
        e[0].record()
        if ivf:
            lo, hi = min(nq, R * per), min(nq, (R + 1) * per)
            annb200._check(L.annb_ivf_route_dev(index.handle, q_t[lo:hi].data_ptr(), hi - lo, dim, k, args.nprobe, tmp_p.data_ptr(), tmp_n.data_ptr(), pitch, st))
        e[1].record()
        if ivf:
            annb200._check(L.annb_ivf_search_probes_shard_dev(index.handle, q_t.data_ptr(), nq, dim, k, args.nprobe, probes.data_ptr(), nprobes.data_ptr(), pitch,
                                                              ids.data_ptr(), dist.data_ptr(), bound.data_ptr(), st))
        else:
            annb200._check(L.annb_flat_search_shard_dev(index.handle, q_t.data_ptr(), nq, dim, k, ids.data_ptr(), dist.data_ptr(), bound.data_ptr(), st))
        e[2].record()
        gathered.view(W, block)[R].copy_(mine)     # (the all-gather's place; other slots stay empty = all-sentinel blocks of zeros are NOT valid,
        if it == -args.warmup:                      #  so fill them once with this rank's block: the merge then sees W equal shards)
            for r in range(W):
                gathered.view(W, block)[r].copy_(mine)
        annb200._check(L.annb_merge_shards_dev(gathered.data_ptr(), block, nq * k * 8, W, nq, k, out_ids.data_ptr(), out_dist.data_ptr(), None, st))
        e[3].record()
        annb200._check(L.annb_shard_check_gathered_dev(index.handle, gathered.data_ptr(), block, nq * k * 12, W, R, out_dist.data_ptr(), nq, k,
                                                       C.byref(mine_c), C.byref(any_c), st))
        e[4].record()
        refine = mine_c.value
        # the deferred step of the serving loop does both in one pass (no read-back)
        annb200._check(L.annb_merge_check_shards_async_dev(index.handle, gathered.data_ptr(), block, nq * k * 8, nq * k * 12, W, R, nq, k, out_ids.data_ptr(),
                                                           out_dist.data_ptr(), h_verdict.data_ptr(), st))
        e[5].record()
    torch.cuda.synchronize()
    tot = 0.0
    for j, nm in enumerate(names):
        ms = float(np.median([ev[i][j].elapsed_time(ev[i][j + 1]) for i in range(args.steps)]))
        tot += ms
        print(f"{nm:14s} {ms:8.3f} ms")
    print(f"{'sum':14s} {tot:8.3f} ms   (world {W}, rank {R}; to-refine of this rank after the merged check: {refine})")
    fused = float(np.median([ev[i][4].elapsed_time(ev[i][5]) for i in range(args.steps)]))
    print(f"{'merge+check':14s} {fused:8.3f} ms   fused pass of the deferred step (annb_merge_check_shards_async_dev), instead of the two lines above")
    print(f"exchange payloads: probes {4 * (per * pitch + per + 1) * W / 1e6:.2f} MB gathered, results {block * W / 1e6:.2f} MB gathered")
    for key in ("last_path", "launches", "fallback_queries", "uncertified"):
        try:
            print(key, index.get_stat(key))
        except Exception:
            pass


if __name__ == "__main__":
    main()
