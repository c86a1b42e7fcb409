"""Multi-device handles (annb_flat_create_multi / annb_ivf_create_multi): one index sharded over several
GPUs behind ONE handle, driven from one process.  Results must be bit-identical to the single-device handle and to the CPU
oracle -- including the order of equal distances, which the merge takes from the shard order.  On a one-GPU box the shards
all live on device 0 (an ordinal may be listed more than once); with more GPUs visible the same tests spread over them."""
import numpy as np
import pytest

import annb200
from annb200 import datagen
from oracle import oracle as o
from util import assert_exact, assert_tie_classes

pytestmark = pytest.mark.gpu

DT = {"f32": (annb200.F32, o.F32), "bf16": (annb200.BF16, o.BF16), "sq8": (annb200.SQ8, o.SQ8)}
MET = {"l2": (annb200.L2, o.L2), "cosine": (annb200.COSINE, o.COSINE)}


def _devices(gpu, shards):
    return [i % gpu for i in range(shards)]


@pytest.mark.parametrize("dtype", ["f32", "bf16", "sq8"])
@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("shards", [1, 3, 8])
def test_flat_multi_equals_single_and_oracle(gpu, dtype, metric, shards):
    data = datagen.gaussian_noise(40_000, 64, seed=3)
    q = datagen.subsample_with_noise(data, 500, seed=3)
    m = annb200.ExhaustiveIndexB200.new(data, MET[metric][0], DT[dtype][0], device=_devices(gpu, shards))
    assert m.shard_count == shards and m.n == 40_000
    ids, d, cnt = m.query_batch(q, 10)
    c = o.build_flat(data, MET[metric][1], DT[dtype][1])
    ref = o.flat_search(c, q, 10)
    assert_exact(ids, d, ref[0], ref[1], f"multi flat {dtype} {metric} x{shards}")
    assert (cnt == 10).all()
    s = annb200.ExhaustiveIndexB200.new(data, MET[metric][0], DT[dtype][0])
    ids1, d1, _ = s.query_batch(q, 10)
    assert_exact(ids, d, ids1, d1, "multi vs single handle")


def test_flat_multi_ties_follow_the_unsharded_order(gpu):
    """Coarse integer data: many equal distances across shard borders.  The merged order must be (distance, row)."""
    rng = np.random.default_rng(5)
    data = rng.integers(-2, 3, size=(20_000, 16)).astype(np.float32)
    q = rng.integers(-2, 3, size=(200, 16)).astype(np.float32)
    m = annb200.ExhaustiveIndexB200.new(data, annb200.L2, annb200.F32, device=_devices(gpu, 5))
    ids, d, _ = m.query_batch(q, 25)
    ref = o.flat_search(o.build_flat(data, o.L2), q, 25)
    assert_exact(ids, d, ref[0], ref[1], "ties across shards")


@pytest.mark.parametrize("dtype", ["f32", "bf16", "sq8"])
def test_flat_multi_self_knn(gpu, dtype):
    """generate_knn over a multi-device handle: query rows come from the shards that own them (peer copies)."""
    data = datagen.correlated(30_000, 50, seed=9)
    m = annb200.ExhaustiveIndexB200.new(data, annb200.L2, DT[dtype][0], device=_devices(gpu, 4))
    ids, d, _ = m.generate_knn(15, row_begin=7_000, row_end=9_000)       # straddles the border between shards 0 and 1
    c = o.build_flat(data, o.L2, DT[dtype][1])
    ref = o.flat_search(c, None, 15, self_rows=np.arange(7_000, 9_000), self_mode=True)
    assert_exact(ids, d, ref[0], ref[1], f"multi self {dtype}")


@pytest.mark.parametrize("dtype", ["f32", "bf16", "sq8"])
@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("shards", [2, 8])
def test_ivf_multi_equals_oracle(gpu, dtype, metric, shards):
    data = datagen.gaussian_noise(60_000, 64, seed=13)
    q = datagen.subsample_with_noise(data, 700, seed=13)
    ci = o.build_ivf(data, MET[metric][1], nlist=128, dtype=DT[dtype][1], kmeans_iters=4)
    norms = ci.norms_i if ci.dtype == o.SQ8 else ci.norms
    m = annb200.IvfIndexB200.from_parts(ci.vectors, ci.centroids, ci.offsets, ci.original_ids, ci.dtype, ci.metric, norms=norms,
                                        centroid_norms=ci.centroid_norms, sq8_scales=ci.scales, device=_devices(gpu, shards))
    assert m.shard_count == shards
    for nprobe in (8, 0):
        ids, d, cnt = m.query_batch(q, 10, nprobe=nprobe or None)
        ref = o.ivf_search(ci, q, 10, nprobe=nprobe or None)
        if dtype == "sq8":
            assert_tie_classes(ids, d, ref[0], ref[1], f"multi ivf sq8 {metric} x{shards}")
        else:
            assert_exact(ids, d, ref[0], ref[1], f"multi ivf {dtype} {metric} x{shards} nprobe={nprobe}")


def test_ivf_multi_ties_follow_list_position(gpu):
    """Equal distances that straddle shards: the unsharded IVF order is (distance, list position), NOT (distance, id)."""
    rng = np.random.default_rng(21)
    base = rng.integers(-2, 3, size=(400, 8)).astype(np.float32)
    data = np.repeat(base, 30, axis=0)                                  # 30 copies of every vector, original ids interleaved below
    perm = rng.permutation(data.shape[0])
    data = np.ascontiguousarray(data[perm])
    q = base[:100] + 0.01
    ci = o.build_ivf(data, o.L2, nlist=32, kmeans_iters=3)
    m = annb200.IvfIndexB200.from_parts(ci.vectors, ci.centroids, ci.offsets, ci.original_ids, ci.dtype, ci.metric, device=_devices(gpu, 4))
    ids, d, _ = m.query_batch(q, 20, nprobe=32)
    ref = o.ivf_search(ci, q, 20, nprobe=32)
    assert_exact(ids, d, ref[0], ref[1], "ivf ties across shards")


def test_multi_handle_errors_and_info(gpu):
    data = datagen.gaussian_noise(5_000, 32, seed=1)
    with pytest.raises(annb200.AnnSearchError) as e:
        annb200.ExhaustiveIndexB200.new(data, annb200.L2, annb200.F32, device=[0, 99])
    assert e.value.variant == "InvalidArgument"
    with pytest.raises(annb200.AnnSearchError) as e:
        annb200.ExhaustiveIndexB200.new(data, annb200.MANHATTAN, annb200.F32, device=[0, 0])
    assert e.value.variant == "DistanceNotSupported"
    m = annb200.ExhaustiveIndexB200.new(data, annb200.L2, annb200.F32, device=[0, 0])
    with pytest.raises(annb200.AnnSearchError) as e:
        m.query_batch(np.zeros((3, 31), np.float32), 5)
    assert e.value.variant == "DimensionMismatch"
    info = m.info()
    assert info.n == 5_000 and info.dim == 32 and info.device == 0 and info.device_bytes >= 5_000 * 32 * 4
    m.set_option("path", annb200.PATH_SIMT)
    ids, d, _ = m.query_batch(data[:10], 3)
    assert (ids[:, 0] == np.arange(10)).all() and m.get_stat("last_path") == annb200.PATH_SIMT


@pytest.mark.parametrize("kind", ["flat", "ivf"])
def test_shard_mode_certificate_against_the_merged_result(gpu, kind):
    """The call sequence of a sharded deployment on ONE device (shards searched one after the other, which is what each
    rank does concurrently): shard-mode searches report a bound instead of certifying locally, the bound is tested against
    the MERGED k-th distances (annb_shard_check_dev), listed queries are recomputed exactly (annb_shard_refine_dev) and
    merged again.  Result: the oracle's rows, and far fewer exact recomputations than local certification would need."""
    import ctypes as C
    import torch
    lib = annb200.lib()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    data = datagen.correlated(120_000, 64, seed=31)
    q = datagen.subsample_with_noise(data, 2000, seed=31)
    nq, dim, k, shards = q.shape[0], 64, 10, 4
    dq = torch.from_numpy(q).to(dev)
    if kind == "flat":
        ref = o.flat_search(o.build_flat(data, o.L2), q, k)
        handles = []
        for s in range(shards):
            lo, hi = (s * data.shape[0]) // shards, ((s + 1) * data.shape[0]) // shards
            handles.append(annb200.ExhaustiveIndexB200.new(data[lo:hi], annb200.L2, annb200.F32, id_base=lo))
        probes = nprobes = None
        pitch, nprobe = 0, 0
    else:
        nprobe = 16
        ci = o.build_ivf(data, o.L2, nlist=256, kmeans_iters=4)
        ref = o.ivf_search(ci, q, k, nprobe=nprobe)
        from annb200 import distributed as D
        handles = []
        for (lb, le) in D.list_ranges(ci.offsets, shards):
            r0, r1 = int(ci.offsets[lb]), int(ci.offsets[le])
            handles.append(annb200.IvfIndexB200.from_parts(ci.vectors[r0:r1], ci.centroids, ci.offsets, ci.original_ids[r0:r1], ci.dtype, ci.metric,
                                                           list_begin=lb, list_end=le, n_total=ci.n))
        pitch = D.probe_pitch(nprobe)
        probes = torch.empty((nq, pitch), dtype=torch.int32, device=dev)
        nprobes = torch.empty((nq,), dtype=torch.int32, device=dev)
        annb200._check(lib.annb_ivf_route_dev(handles[0].handle, dq.data_ptr(), nq, dim, k, nprobe, probes.data_ptr(), nprobes.data_ptr(), pitch, st))
    block = ((nq * k * 12 + 255) // 256) * 256
    gathered = torch.empty((shards * block,), dtype=torch.uint8, device=dev)
    bounds = [torch.empty((nq,), dtype=torch.float32, device=dev) for _ in range(shards)]
    m_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    m_d = torch.empty((nq, k), dtype=torch.float32, device=dev)

    def slot(s):
        b = gathered[s * block:(s + 1) * block]
        return b.data_ptr(), b[nq * k * 8:].data_ptr()

    for s, h in enumerate(handles):
        h.set_option("path", annb200.PATH_TENSOR if kind == "flat" else annb200.PATH_AUTO)
        ids_p, d_p = slot(s)
        if kind == "flat":
            annb200._check(lib.annb_flat_search_shard_dev(h.handle, dq.data_ptr(), nq, dim, k, ids_p, d_p, bounds[s].data_ptr(), st))
        else:
            annb200._check(lib.annb_ivf_search_probes_shard_dev(h.handle, dq.data_ptr(), nq, dim, k, nprobe, probes.data_ptr(), nprobes.data_ptr(), pitch,
                                                                ids_p, d_p, bounds[s].data_ptr(), st))
    annb200._check(lib.annb_merge_shards_dev(gathered.data_ptr(), block, nq * k * 8, shards, nq, k, m_ids.data_ptr(), m_d.data_ptr(), None, st))
    refined = 0
    for s, h in enumerate(handles):
        cnt = C.c_uint32(0)
        annb200._check(lib.annb_shard_check_dev(h.handle, bounds[s].data_ptr(), m_d.data_ptr(), nq, k, C.byref(cnt), st))
        refined += cnt.value
        if cnt.value:
            ids_p, d_p = slot(s)
            annb200._check(lib.annb_shard_refine_dev(h.handle, dq.data_ptr(), nq, dim, k, nprobe, None if probes is None else probes.data_ptr(),
                                                     None if nprobes is None else nprobes.data_ptr(), pitch, ids_p, d_p, st))
    if refined:
        annb200._check(lib.annb_merge_shards_dev(gathered.data_ptr(), block, nq * k * 8, shards, nq, k, m_ids.data_ptr(), m_d.data_ptr(), None, st))
    torch.cuda.synchronize()
    assert_exact(m_ids.cpu().numpy(), m_d.cpu().numpy(), ref[0], ref[1], f"shard mode {kind}")
    # local certification on the same shards, for comparison: how many queries each shard would have recomputed on its own
    local = 0
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    for s, h in enumerate(handles):
        f0 = h.get_stat("fallback_queries")
        if kind == "flat":
            annb200._check(lib.annb_flat_search_dev(h.handle, dq.data_ptr(), nq, dim, k, out_i.data_ptr(), None, None, st))
        else:
            annb200._check(lib.annb_ivf_search_probes_dev(h.handle, dq.data_ptr(), nq, dim, k, nprobe, probes.data_ptr(), nprobes.data_ptr(), pitch,
                                                          out_i.data_ptr(), None, None, st))
        torch.cuda.synchronize()
        local += h.get_stat("fallback_queries") - f0
    print(f"\\n[shard mode {kind}] queries refined after the merged check: {refined}; recomputed under local certification: {local}")
    assert refined <= local


@pytest.mark.parametrize("parts,k", [(1, 10), (3, 10), (8, 15), (8, 16), (12, 10), (9, 15)])
def test_fused_merge_and_check_equals_the_two_passes(gpu, parts, k):
    """annb_merge_check_shards_async_dev = annb_merge_shards_dev + annb_shard_check_gathered_async_dev on the same gathered blocks:
    same merged rows (ties across shards, short lists, empty shards) and the same verdict words (bounds that fail, infinite
    bounds, a shard whose status word is set).  parts * k <= 128 takes the fused kernel, larger merges the two passes inside."""
    import torch
    lib = annb200.lib()
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(parts * 100 + k)
    nq = 777
    data = datagen.gaussian_noise(5000, 16, seed=3)
    h = annb200.ExhaustiveIndexB200.new(data, annb200.L2, annb200.F32)      # any handle: the call only uses its scratch and ordering
    block = ((nq * k * 12 + nq * 4 + 4 + 255) // 256) * 256
    host = np.zeros((parts, block), np.uint8)
    for s_ in range(parts):
        d = np.sort(rng.integers(0, 40, (nq, k)).astype(np.float32) * 0.5, axis=1)      # coarse values: many ties across shards
        ids = (rng.integers(0, 1 << 40, (nq, k), dtype=np.int64)).astype(np.uint64)
        valid = rng.integers(0, k + 1, nq) if s_ % 3 == 1 else np.full(nq, k)          # some shards hold short (or empty) lists
        for q_ in np.nonzero(valid < k)[0]:
            ids[q_, valid[q_]:] = np.uint64(0xFFFFFFFFFFFFFFFF)
            d[q_, valid[q_]:] = np.inf
        bound = rng.choice(np.array([np.inf, 1.0, 5.0, 12.0, 30.0], np.float32), nq)
        host[s_, :nq * k * 8] = ids.view(np.uint8).reshape(-1)
        host[s_, nq * k * 8:nq * k * 12] = d.view(np.uint8).reshape(-1)
        host[s_, nq * k * 12:nq * k * 12 + nq * 4] = bound.view(np.uint8)
    for status_part in (None, parts - 1):
        if status_part is not None:
            host[status_part, nq * k * 12 + nq * 4:nq * k * 12 + nq * 4 + 4] = np.array([1], np.int32).view(np.uint8)
        gathered = torch.from_numpy(host.reshape(-1)).to(dev)
        for my in sorted({0, parts - 1}):
            a_ids = torch.empty((nq, k), dtype=torch.int64, device=dev); a_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
            b_ids = torch.empty_like(a_ids); b_d = torch.empty_like(a_d)
            va = torch.zeros((2,), dtype=torch.int32).pin_memory(); vb = torch.zeros((2,), dtype=torch.int32).pin_memory()
            annb200._check(lib.annb_merge_shards_dev(gathered.data_ptr(), block, nq * k * 8, parts, nq, k, a_ids.data_ptr(), a_d.data_ptr(), None, st))
            annb200._check(lib.annb_shard_check_gathered_async_dev(h.handle, gathered.data_ptr(), block, nq * k * 12, parts, my, a_d.data_ptr(), nq, k,
                                                                   va.data_ptr(), st))
            annb200._check(lib.annb_merge_check_shards_async_dev(h.handle, gathered.data_ptr(), block, nq * k * 8, nq * k * 12, parts, my, nq, k,
                                                                 b_ids.data_ptr(), b_d.data_ptr(), vb.data_ptr(), st))
            torch.cuda.synchronize()
            assert torch.equal(a_ids, b_ids) and torch.equal(a_d.view(torch.int32), b_d.view(torch.int32)), (parts, k, my)
            assert va.tolist() == vb.tolist(), (parts, k, my, va.tolist(), vb.tolist())
            assert (vb[1].item() & 2) == (0 if status_part is None else 2)
            # the verdict against a plain numpy statement
            dk = a_d.cpu().numpy()[:, k - 1]
            bounds = np.stack([host[s_, nq * k * 12:nq * k * 12 + nq * 4].view(np.float32) for s_ in range(parts)])
            need = ~(bounds > dk[None, :]) & np.isfinite(bounds)
            assert vb[0].item() == int(need[my].sum()) and (vb[1].item() & 1) == int(need.any())


@pytest.mark.parametrize("kind", ["flat", "ivf"])
def test_sharded_search_deferred_verdict(gpu, kind):
    """ShardedSearch on a one-rank NCCL group: a deferred step (verdict read later, resolve()) gives the rows of the immediate
    step and of the oracle, also when the verdict asks for a refine (forced with a pessimistic certificate bound), and two
    objects can take turns on one handle (the serving loop of bench.py)."""
    import socket
    import torch
    import torch.distributed as dist
    from annb200 import distributed as D
    own_group = not dist.is_initialized()
    if own_group:
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        dev = torch.device("cuda:0")
        torch.cuda.set_device(dev)
        data = datagen.correlated(60_000, 64, seed=33)
        q = datagen.subsample_with_noise(data, 1500, seed=33)
        nq, dim, k = q.shape[0], 64, 10
        dq = torch.from_numpy(q).to(dev)
        if kind == "flat":
            ref = o.flat_search(o.build_flat(data, o.L2), q, k)
            h = annb200.ExhaustiveIndexB200.new(data, annb200.L2, annb200.F32)
            nprobe = 0
        else:
            nprobe = 12
            ci = o.build_ivf(data, o.L2, nlist=128, kmeans_iters=4)
            ref = o.ivf_search(ci, q, k, nprobe=nprobe)
            h = annb200.IvfIndexB200.from_parts(ci.vectors, ci.centroids, ci.offsets, ci.original_ids, ci.dtype, ci.metric)
        a, b = D.ShardedSearch(h, nq, dim, k, nprobe, None, dev), D.ShardedSearch(h, nq, dim, k, nprobe, None, dev)
        ids, dd = a(dq)
        torch.cuda.synchronize()
        assert_exact(ids.cpu().numpy(), dd.cpu().numpy(), ref[0], ref[1], f"immediate {kind}")
        for eps in (1, -3):                       # derived bound; absurd bound: (nearly) every query is refined after the merged check
            h.set_option("cert_eps_log2", eps)
            r0 = a.refined_queries + b.refined_queries
            a(dq, defer=True)
            b(dq, defer=True)                     # enqueued before a's verdict is read
            a.resolve()
            b.resolve()
            torch.cuda.synchronize()
            for s in (a, b):
                assert_exact(s.out_ids.cpu().numpy(), s.out_dist.cpu().numpy(), ref[0], ref[1], f"deferred {kind} eps {eps}")
            if eps == -3:
                assert a.refined_queries + b.refined_queries - r0 > nq, "the pessimistic bound should have forced refines on both steps"
    finally:
        if own_group:
            dist.destroy_process_group()
