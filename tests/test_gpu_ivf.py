"""GPU parity: IVF query (centroid ranking -> probe expansion -> list scan -> top-k) through the C ABI against
the CPU oracle on *identical index contents* (same centroids, offsets, permutation)."""
import numpy as np
import pytest

import annb200
from annb200 import datagen
from oracle import oracle as o
from util import assert_exact, assert_tie_classes

pytestmark = pytest.mark.gpu

SIMPLE = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1]], dtype=np.float32)
DT = {"f32": (annb200.F32, o.F32), "bf16": (annb200.BF16, o.BF16), "sq8": (annb200.SQ8, o.SQ8)}
MET = {"l2": (annb200.L2, o.L2), "cosine": (annb200.COSINE, o.COSINE)}


def _gpu_from_oracle(ix: o.IvfIndex, list_begin=0, list_end=None):
    list_end = ix.nlist if list_end is None else list_end
    r0, r1 = int(ix.offsets[list_begin]), int(ix.offsets[list_end])
    norms = ix.norms_i if ix.dtype == o.SQ8 else ix.norms
    return annb200.IvfIndexB200.from_parts(
        ix.vectors[r0:r1], ix.centroids, ix.offsets, ix.original_ids[r0:r1], ix.dtype, ix.metric,
        norms=None if norms is None else norms[r0:r1], centroid_norms=ix.centroid_norms, sq8_scales=ix.scales,
        list_begin=list_begin, list_end=list_end, n_total=ix.n)


def _check(dtype, got, ref, what):
    # IVF-SQ8 keeps a heap filled in probe order: the subset kept inside the last tie class is
    # implementation-defined in the reference (src/quantised/ivf_sq8.rs:329-352)
    (assert_tie_classes if dtype == "sq8" else assert_exact)(got[0], got[1], ref[0], ref[1], what)
    assert np.array_equal(got[2], ref[2]), what + " counts"


def test_probe_expansion_singletons(gpu):
    # src/cpu/ivf.rs:777-803: five singleton cells, nprobe = 1, k = 3 -> expands to 3 cells
    c = o.build_ivf(SIMPLE, o.L2, nlist=5, centroids=SIMPLE.copy())
    g = _gpu_from_oracle(c)
    ids, d, cnt = g.query_batch(np.array([[1, 0, 0]], np.float32), 3, nprobe=1)
    rids, rd, rcnt, npb, nsc = o.ivf_search(c, np.array([[1, 0, 0]], np.float32), 3, nprobe=1)
    assert cnt[0] == 3 and ids[0, 0] == 0
    assert_exact(ids, d, rids, rd, "probe expansion")
    assert g.get_stat("probed_lists") == int(npb.sum()) and g.get_stat("scanned_vectors") == int(nsc.sum())
    ids, d, cnt = g.query_batch(SIMPLE, 10, nprobe=2)    # k > n: every cell, 5 results
    assert (cnt == 5).all() and (ids[:, 5:] == -1).all()


@pytest.mark.parametrize("dtype", ["f32", "bf16", "sq8"])
@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("n,dim,nlist,nq,k,nprobe", [(6000, 32, 64, 200, 10, 8), (2500, 50, 40, 33, 15, None), (3000, 128, 32, 64, 10, 4),
                                                     (900, 7, 30, 17, 31, 1)])
def test_exact_parity_external_queries(gpu, dtype, metric, n, dim, nlist, nq, k, nprobe):
    data = datagen.gaussian_noise(n, dim, seed=21)
    q = datagen.subsample_with_noise(data, nq, seed=21)
    c = o.build_ivf(data, MET[metric][1], nlist=nlist, dtype=DT[dtype][1], kmeans_iters=5)
    g = _gpu_from_oracle(c)
    got = g.query_batch(q, k, nprobe=nprobe)
    ref = o.ivf_search(c, q, k, nprobe=nprobe)
    _check(dtype, got, ref, f"ivf {dtype} {metric} n={n} dim={dim} nprobe={nprobe}")
    assert g.get_stat("scanned_vectors") == int(ref[4].sum())
    g.set_option("ivf_list_major", 1)      # list-major batched scans: (list, query group) tasks
    g.set_option("path", annb200.PATH_SIMT)
    _check(dtype, g.query_batch(q, k, nprobe=nprobe), ref, "list-major CUDA-core scan")
    if k <= 24:                            # tensor-core grouped scan (tcgen05) + exact re-rank
        g.set_option("path", annb200.PATH_TENSOR)
        _check(dtype, g.query_batch(q, k, nprobe=nprobe), ref, "list-major tensor-core scan")
        assert g.get_stat("last_path") == annb200.PATH_TENSOR
    g.set_option("path", annb200.PATH_AUTO)
    g.set_option("ivf_list_major", 0)      # query-major streaming scan: one warp per (query, part)
    for parts in (1, 3):       # result must not depend on how probes are split across warps
        g.set_option("scan_parts", parts)
        _check(dtype, g.query_batch(q, k, nprobe=nprobe), ref, f"query-major scan_parts={parts}")


@pytest.mark.parametrize("dtype", ["f32", "bf16", "sq8"])
@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_exact_parity_self_queries(gpu, dtype, metric):
    data = datagen.correlated(1500, 32, seed=5)
    c = o.build_ivf(data, MET[metric][1], nlist=24, dtype=DT[dtype][1], kmeans_iters=5)
    g = _gpu_from_oracle(c)
    got = g.generate_knn(10, nprobe=6)                        # scattered to original ids (ivf.rs:476-486)
    ref = o.ivf_search(c, None, 10, nprobe=6, self_mode=True)
    _check(dtype, got, ref, f"ivf self {dtype} {metric}")
    got = g.generate_knn(4, nprobe=3, pos_begin=10, pos_end=100)
    ref = o.ivf_search(c, None, 4, nprobe=3, self_rows=np.arange(10, 100), self_mode=True)
    _check(dtype, got, ref, f"ivf self sub-range {dtype} {metric}")


def test_empty_lists_and_imbalance(gpu):
    # cells 1 and 3 are empty; cell 0 holds most of the data
    rng = np.random.default_rng(2)
    data = np.concatenate([rng.standard_normal((900, 16)) * 0.1, rng.standard_normal((100, 16)) * 0.1 + 5]).astype(np.float32)
    cent = np.stack([np.zeros(16), np.full(16, 50.0), np.full(16, 5.0), np.full(16, -50.0)]).astype(np.float32)
    c = o.build_ivf(data, o.L2, nlist=4, centroids=cent)
    assert c.offsets[2] == c.offsets[1] and c.offsets[4] == c.offsets[3]
    g = _gpu_from_oracle(c)
    q = np.full((3, 16), 49.0, np.float32)        # nearest centroid is an empty cell
    got = g.query_batch(q, 5, nprobe=1)
    ref = o.ivf_search(c, q, 5, nprobe=1)
    _check("f32", got, ref, "empty nearest cell")


def test_sharded_lists_merge_to_the_unsharded_answer(gpu):
    """Multi-GPU layout on one device: two handles own disjoint list ranges; per-shard top-k merged with
    annb_merge_topk_dev equals the single-index answer."""
    import ctypes as C
    import torch
    data = datagen.gaussian_noise(5000, 32, seed=31)
    q = datagen.subsample_with_noise(data, 128, seed=31)
    c = o.build_ivf(data, o.L2, nlist=32, kmeans_iters=5)
    k, nprobe = 10, 6
    ref = o.ivf_search(c, q, k, nprobe=nprobe)
    shards = [_gpu_from_oracle(c, 0, 13), _gpu_from_oracle(c, 13, 32)]
    dq = torch.from_numpy(q).cuda()
    pid = torch.empty((2, q.shape[0], k), dtype=torch.int64, device="cuda")
    pd = torch.empty((2, q.shape[0], k), dtype=torch.float32, device="cuda")
    lib = annb200.lib()
    st = torch.cuda.current_stream().cuda_stream
    for i, s in enumerate(shards):
        annb200._check(lib.annb_ivf_search_dev(s.handle, dq.data_ptr(), q.shape[0], q.shape[1], k, nprobe, pid[i].data_ptr(), pd[i].data_ptr(), None, st))
    oid = torch.empty((q.shape[0], k), dtype=torch.int64, device="cuda")
    od = torch.empty((q.shape[0], k), dtype=torch.float32, device="cuda")
    oc = torch.empty((q.shape[0],), dtype=torch.int32, device="cuda")
    annb200._check(lib.annb_merge_topk_dev(pid.data_ptr(), pd.data_ptr(), 2, q.shape[0], k, oid.data_ptr(), od.data_ptr(), oc.data_ptr(), st))
    torch.cuda.synchronize()
    ids, d = oid.cpu().numpy(), od.cpu().numpy()
    # merged order is (distance, original id); the unsharded order is (distance, list position): compare as sets per distance
    assert (d.view(np.uint32) == ref[1].view(np.uint32)).all()
    assert (np.sort(ids, axis=1) == np.sort(ref[0], axis=1)).all()


def test_coarse_assignment_matches_direct_assign(gpu):
    # src/gpu/k_means_gpu.rs:2887-2924: assign_all_gpu == assign_all_parallel exactly
    for metric in ("l2", "cosine"):
        for dim in (2, 8, 50, 128):
            data = datagen.gaussian_noise(3000, dim, seed=13)
            cent = data[::100].copy()
            cn = np.array([o.seq_norm_f32(r) for r in cent], np.float32)
            a = annb200.ivf_assign(data, cent, MET[metric][0], cn if metric == "cosine" else None)
            r = o.assign_all(data, cent, cn, MET[metric][1])
            assert np.array_equal(a.astype(np.int64), r), f"{metric} dim={dim}"


def _assign_redone():
    import ctypes as C
    f = annb200.lib().annb_assign_last_redone
    f.restype = C.c_uint64
    return int(f())


@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("dim", [32, 50, 128])
def test_tensor_core_assignment_matches_direct_assign(gpu, metric, dim):
    """Tables of >= 512 centroids: annb_ivf_assign pre-selects on the tensor cores, recomputes the kept cells' scores in the
    direct_assign arithmetic (src/utils/k_means_utils.rs:2119-2195) and certifies the winner -- the assignments must be the
    oracle's, bit for bit, including duplicate centroids (lowest cell wins the tie) and a zero centroid."""
    data = datagen.gaussian_noise(20000, dim, seed=71)
    rng = np.random.default_rng(7)
    cent = data[rng.choice(20000, 1000, replace=False)].copy()
    cent[500:520] = cent[100:120]          # exact duplicates: ties broken by the lower cell
    cent[777] = 0.0                        # zero centroid: cosine score 0 (1/|c| := 0), L2 score -0
    cn = np.array([o.seq_norm_f32(r) for r in cent], np.float32)
    a = annb200.ivf_assign(data, cent, MET[metric][0], cn if metric == "cosine" else None)
    redone = _assign_redone()
    r = o.assign_all(data, cent, cn, MET[metric][1])
    assert np.array_equal(a.astype(np.int64), r), f"{metric} dim={dim}: {(a.astype(np.int64) != r).sum()} rows differ"
    assert redone < 200, redone            # the certificate passes for nearly every row
    assert not np.isin(a, np.arange(500, 520)).any()


def test_tensor_core_assignment_uncertifiable_rows_are_redone(gpu):
    """Rows that coincide with many identical centroids cannot be separated from the pruning threshold: they are redone on
    the exact kernel and still get the lowest cell."""
    data = datagen.gaussian_noise(8000, 32, seed=73)
    cent = np.concatenate([np.repeat(data[:1], 40, axis=0), data[100:700]]).astype(np.float32)   # cell 0..39 identical
    pts = np.concatenate([np.repeat(data[:1], 300, axis=0), data[1000:8000]]).astype(np.float32)
    cn = np.array([o.seq_norm_f32(r) for r in cent], np.float32)
    a = annb200.ivf_assign(pts, cent, annb200.L2, None)
    assert _assign_redone() >= 300
    r = o.assign_all(pts, cent, cn, o.L2)
    assert np.array_equal(a.astype(np.int64), r)
    assert (a[:300] == 0).all()


def test_device_lloyd_on_the_tensor_path(gpu):
    """annb_kmeans_lloyd with a table large enough for the tensor-core assignment: same iterations and centroids as the
    restated parallel_lloyd (src/utils/k_means_utils.rs:1572-1700)."""
    data = datagen.gaussian_noise(12000, 32, seed=77)
    rng = np.random.default_rng(9)
    init = data[rng.choice(12000, 600, replace=False)].copy()
    c, it = annb200.kmeans_lloyd(data, init, annb200.L2, max_iters=6)
    r, rit = o.parallel_lloyd(data, init, o.L2, max_iters=6)
    assert it == rit
    np.testing.assert_allclose(c, r, rtol=1e-5, atol=1e-5)


def test_mirror_build_and_query(gpu):
    data = datagen.gaussian_noise(4000, 32, seed=17)
    q = datagen.subsample_with_noise(data, 50, seed=17)
    flat = annb200.build_exhaustive_index_gpu(data, "euclidean")
    truth, _ = annb200.query_exhaustive_index_gpu(q, flat, 10)
    for build in (annb200.build_ivf_index_gpu, annb200.build_ivf_bf16_index, annb200.build_ivf_sq8_index):
        ix = build(data, nlist=32, k_means_params={"iters": 5}, dist_metric="euclidean", seed=42)
        ids, dist = annb200.query_ivf_index_gpu(q, ix, 10, nprobe=32)
        rec = o.recall_at_k(truth, ids, 10)
        assert rec > (0.6 if build is annb200.build_ivf_sq8_index else 0.95), (build.__name__, rec)
    with pytest.raises(annb200.AnnSearchError) as e:
        annb200.build_ivf_index_gpu(data, nlist=8, dist_metric="manhattan")
    assert e.value.variant == "DistanceNotSupported"


@pytest.mark.parametrize("dtype", ["f32", "sq8"])
def test_fast_probe_path_and_overflow_fallback(gpu, dtype):
    """nlist > 1024 ranks only nprobe + 64 centroids with the fused select; a query that needs more cells than that to
    reach k vectors (tiny lists) must fall back to the full ranking -- both must equal the oracle."""
    data = datagen.gaussian_noise(6000, 24, seed=41)
    q = datagen.subsample_with_noise(data, 96, seed=41)
    c = o.build_ivf(data, o.L2, nlist=1500, dtype=DT[dtype][1], kmeans_iters=2)      # ~4 vectors per list
    g = _gpu_from_oracle(c)
    for k, nprobe in ((10, 8), (10, 40), (200, 1), (5, 1499)):     # (200, 1): needs ~50+ cells; may or may not overflow
        ref = o.ivf_search(c, q, k, nprobe=nprobe)
        for tc_coarse, fast in ((0, 2), (0, 0), (1, 1)):      # fast = 2: force the fused select regardless of the prefix length
            g.set_option("ivf_tc_coarse", tc_coarse)
            g.set_option("ivf_fast_probe", fast)
            _check(dtype, g.query_batch(q, k, nprobe=nprobe), ref, f"tc_coarse={tc_coarse} fast_probe={fast} k={k} nprobe={nprobe}")
            assert g.get_stat("scanned_vectors") == int(ref[4].sum())
            if tc_coarse == 0:
                assert g.get_stat("coarse_path") != 2
    ref = o.ivf_search(c, q, 600, nprobe=1)                        # certainly more than nprobe + 64 cells
    for tc_coarse in (0, 1):
        g.set_option("ivf_tc_coarse", tc_coarse)
        g.set_option("ivf_fast_probe", 2)
        _check(dtype, g.query_batch(q, 600, nprobe=1), ref, "overflow fallback")
        assert g.get_stat("coarse_path") == 0


def test_ivf_tensor_certificate_fallback(gpu):
    data = datagen.gaussian_noise(8000, 32, seed=33)
    q = datagen.subsample_with_noise(data, 150, seed=33)
    c = o.build_ivf(data, o.L2, nlist=40, kmeans_iters=4)
    g = _gpu_from_oracle(c)
    g.set_option("ivf_list_major", 1)
    g.set_option("path", annb200.PATH_TENSOR)
    ref = o.ivf_search(c, q, 10, nprobe=6)
    g.set_option("cert_eps_log2", -2)          # every query fails the certificate -> exact pipeline for all of them
    got = g.query_batch(q, 10, nprobe=6)
    assert g.get_stat("fallback_queries") == 150
    _check("f32", got, ref, "ivf fallback")
    assert g.get_stat("scanned_vectors") == int(ref[4].sum())
    g.set_option("cert_eps_log2", -20)
    _check("f32", g.query_batch(q, 10, nprobe=6), ref, "ivf default bound")


@pytest.mark.parametrize("dtype", ["f32", "bf16", "sq8"])
@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_tensor_core_centroid_ranking(gpu, dtype, metric):
    """nlist >= 512: the centroid ranking runs on the tensor cores (dense 3xTF32 values -> radix select -> exact distances
    of the candidates -> certified prefix).  The probe sets, and therefore the results and the scan statistics, must be
    those of the exact ranking."""
    data = datagen.gaussian_noise(30000, 32, seed=51)
    q = datagen.subsample_with_noise(data, 300, seed=51)
    c = o.build_ivf(data, MET[metric][1], nlist=700, dtype=DT[dtype][1], kmeans_iters=3)
    g = _gpu_from_oracle(c)
    for k, nprobe in ((10, 8), (10, 60), (24, 1)):
        ref = o.ivf_search(c, q, k, nprobe=nprobe)
        got = g.query_batch(q, k, nprobe=nprobe)
        assert g.get_stat("coarse_path") == 2
        _check(dtype, got, ref, f"tc coarse {dtype} {metric} k={k} nprobe={nprobe}")
        assert g.get_stat("scanned_vectors") == int(ref[4].sum()) and g.get_stat("probed_lists") == int(ref[3].sum())
    got = g.generate_knn(10, nprobe=12, pos_begin=100, pos_end=700)        # self queries: routing on decoded rows
    assert g.get_stat("coarse_path") == 2
    ref = o.ivf_search(c, None, 10, nprobe=12, self_rows=np.arange(100, 700), self_mode=True)
    _check(dtype, got, ref, f"tc coarse self {dtype} {metric}")


def test_tensor_core_centroid_ranking_ties_and_uncertifiable(gpu):
    """Duplicate centroids tie exactly (broken by cell id, as the exact ranking does); a tie class larger than the
    candidate capacity, or an error bound that certifies nothing, sends the batch to the exact ranking."""
    rng = np.random.default_rng(5)
    data = datagen.gaussian_noise(20000, 16, seed=61)
    base = data[rng.choice(20000, 200, replace=False)]
    cent = np.concatenate([base, base, base, base[:40]]).astype(np.float32)       # 640 cells, every centre 3-4 times
    c = o.build_ivf(data, o.L2, nlist=640, centroids=cent)
    g = _gpu_from_oracle(c)
    q = datagen.subsample_with_noise(data, 200, seed=61)
    ref = o.ivf_search(c, q, 10, nprobe=9)
    _check("f32", g.query_batch(q, 10, nprobe=9), ref, "duplicate centroids")
    assert g.get_stat("coarse_path") == 2
    g.set_option("cert_eps_log2", -2)                 # nothing certifiable -> exact ranking, same answer
    _check("f32", g.query_batch(q, 10, nprobe=9), ref, "uncertifiable ranking")
    assert g.get_stat("coarse_path") != 2
    g.set_option("cert_eps_log2", -20)
    cent2 = np.concatenate([np.repeat(base[:1], 300, axis=0), base, base[:140]]).astype(np.float32)   # one centre 301 times
    c2 = o.build_ivf(data, o.L2, nlist=640, centroids=cent2)
    g2 = _gpu_from_oracle(c2)
    q2 = (base[:1] + 0.01 * rng.standard_normal((40, 16))).astype(np.float32)
    ref2 = o.ivf_search(c2, q2, 10, nprobe=9)
    _check("f32", g2.query_batch(q2, 10, nprobe=9), ref2, "tie class beyond the candidate capacity")
    assert g2.get_stat("coarse_path") != 2


@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("dim", [16, 50])
def test_device_lloyd_matches_the_restated_parallel_lloyd(gpu, metric, dim):
    """annb_kmeans_lloyd vs the oracle's restatement of parallel_lloyd (k_means_utils.rs:1572-1700) from the same
    initial centroids.  One step: assignments are bit-exact (direct_assign arithmetic), the means agree to f32
    rounding of an f64 sum; a full run stops after the same number of updates with centroids within 1e-5."""
    data = datagen.gaussian_noise(6000, dim, seed=71)
    rng = np.random.default_rng(3)
    init = data[rng.choice(6000, 48, replace=False)].copy()
    m_g, m_o = MET[metric]
    c1, it1 = annb200.kmeans_lloyd(data, init, m_g, max_iters=1)
    r1, rit1 = o.parallel_lloyd(data, init, m_o, max_iters=1)
    assert it1 == rit1 == 1
    np.testing.assert_allclose(c1, r1, rtol=2e-7, atol=1e-7)
    c, it = annb200.kmeans_lloyd(data, init, m_g, max_iters=30)
    r, rit = o.parallel_lloyd(data, init, m_o, max_iters=30)
    assert it == rit and 1 < it <= 30
    np.testing.assert_allclose(c, r, rtol=1e-5, atol=1e-5)
    # the trained centroids are a fixed point up to the stop rule: one more assignment changes few points
    cn = np.array([o.l2_norm_f32(x) for x in c], np.float32)
    a = annb200.ivf_assign(data, c, m_g, cn if metric == "cosine" else None)
    assert len(np.unique(a)) > 40


def test_device_lloyd_edge_cases(gpu):
    data = datagen.gaussian_noise(500, 8, seed=5)
    # an initial centroid far from every point stays where it is (empty clusters keep their centroid)
    init = np.concatenate([data[:3], np.full((1, 8), 1e3, np.float32)])
    c, it = annb200.kmeans_lloyd(data, init, annb200.L2, max_iters=5)
    assert np.array_equal(c[3], init[3]) and it >= 1
    r, rit = o.parallel_lloyd(data, init, o.L2, max_iters=5)
    assert it == rit
    np.testing.assert_allclose(c, r, rtol=1e-5, atol=1e-5)
    with pytest.raises(annb200.AnnSearchError) as e:          # TooFewSamplesForCentroids (k_means_utils.rs:2788-2793)
        annb200.kmeans_lloyd(data[:3], data[:4], annb200.L2)
    assert e.value.variant == "TooFewSamplesForCentroids"
    with pytest.raises(annb200.AnnSearchError) as e:
        annb200.kmeans_lloyd(data, data[:4], annb200.MANHATTAN)
    assert e.value.variant == "DistanceNotSupported"


@pytest.mark.parametrize("dtype", ["f32", "sq8"])
def test_route_then_search_with_probes_equals_the_one_call_search(gpu, dtype):
    """annb_ivf_route_dev + annb_ivf_search_probes_dev are the two halves of annb_ivf_search_dev.  Emulates two ranks on
    one device: each "rank" routes its half of the batch, the probe lists are concatenated, each shard scans its own
    lists for the whole batch with those lists, and the merged answer equals the unsharded search."""
    import torch
    from annb200 import distributed as D
    data = datagen.gaussian_noise(30000, 32, seed=81)
    q = datagen.subsample_with_noise(data, 257, seed=81)
    c = o.build_ivf(data, o.L2, nlist=600, dtype=DT[dtype][1], kmeans_iters=3)
    k, nprobe = 10, 12
    ref = o.ivf_search(c, q, k, nprobe=nprobe)
    full = _gpu_from_oracle(c)
    lb = D.list_ranges(c.offsets, 2)
    shards = [_gpu_from_oracle(c, *lb[0]), _gpu_from_oracle(c, *lb[1])]
    lib = annb200.lib()
    dq = torch.from_numpy(q).cuda()
    st = torch.cuda.current_stream().cuda_stream
    pitch = D.probe_pitch(nprobe)
    nq = q.shape[0]
    probes = torch.full((nq, pitch), -1, dtype=torch.int32, device="cuda")
    npb = torch.zeros((nq,), dtype=torch.int32, device="cuda")
    half = 129
    for r, (lo, hi) in enumerate(((0, half), (half, nq))):      # rank r routes its slice on its own shard handle
        annb200._check(lib.annb_ivf_route_dev(shards[r].handle, dq[lo:hi].data_ptr(), hi - lo, 32, k, nprobe, probes[lo:hi].data_ptr(),
                                              npb[lo:hi].data_ptr(), pitch, st))
    torch.cuda.synchronize()
    assert int(npb.sum().item()) == int(ref[3].sum())            # same probe counts as select_probed_clusters
    # one handle, whole index: halves == one call
    ids = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    dd = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    cnt = torch.empty((nq,), dtype=torch.int32, device="cuda")
    annb200._check(lib.annb_ivf_search_probes_dev(full.handle, dq.data_ptr(), nq, 32, k, nprobe, probes.data_ptr(), npb.data_ptr(), pitch,
                                                  ids.data_ptr(), dd.data_ptr(), cnt.data_ptr(), st))
    torch.cuda.synchronize()
    _check(dtype, (ids.cpu().numpy(), dd.cpu().numpy(), cnt.cpu().numpy().astype(ref[2].dtype)), ref, "route + search_probes")
    # two shards: scan own lists for the whole batch, merge
    pid = torch.empty((2, nq, k), dtype=torch.int64, device="cuda")
    pd = torch.empty((2, nq, k), dtype=torch.float32, device="cuda")
    for r in range(2):
        annb200._check(lib.annb_ivf_search_probes_dev(shards[r].handle, dq.data_ptr(), nq, 32, k, nprobe, probes.data_ptr(), npb.data_ptr(), pitch,
                                                      pid[r].data_ptr(), pd[r].data_ptr(), None, st))
    m_ids, m_d = D.merge_topk_device(pid, pd)
    torch.cuda.synchronize()
    assert (m_d.cpu().numpy().view(np.uint32) == ref[1].view(np.uint32)).all()
    if dtype != "sq8":
        assert (np.sort(m_ids.cpu().numpy(), axis=1) == np.sort(ref[0], axis=1)).all()
    # a pitch that cannot hold the expanded probe set is reported, not truncated
    small = torch.empty((nq, 4), dtype=torch.int32, device="cuda")
    rc = lib.annb_ivf_route_dev(full.handle, dq.data_ptr(), nq, 32, k, nprobe, small.data_ptr(), npb.data_ptr(), 4, st)
    assert rc == -8


@pytest.mark.parametrize("nq", [1, 3, 16, 40])
def test_small_batches_both_scan_kernels(gpu, nq):
    """Small batches: the query-major streaming kernel and the list-major kernels (forced through `ivf_list_major`) and the
    library's own choice all return the oracle's answer (src/cpu/ivf.rs:337-390)."""
    data = datagen.gaussian_noise(30000, 32, seed=23)
    c = o.build_ivf(data, o.L2, nlist=64)
    g = _gpu_from_oracle(c)
    q = datagen.subsample_with_noise(data, nq, seed=29)
    ref = o.ivf_search(c, q, 10, nprobe=6)
    for mode in (0, 1, -1):
        g.set_option("ivf_list_major", mode)
        _check("f32", g.query_batch(q, 10, nprobe=6), ref, f"nq={nq} ivf_list_major={mode}")


def test_concurrent_ivf_searches_on_one_handle(gpu):
    """Several host threads searching one IVF index at once (`&self` queries, src/cpu/ivf.rs:337) all get the oracle's answer."""
    import threading
    data = datagen.gaussian_noise(30000, 32, seed=31)
    c = o.build_ivf(data, o.L2, nlist=64)
    g = _gpu_from_oracle(c)
    qs = [datagen.subsample_with_noise(data, 200 + 64 * t, seed=200 + t) for t in range(4)]
    refs = [o.ivf_search(c, q, 10, nprobe=5 + t) for t, q in enumerate(qs)]
    out, errs = [None] * 4, []

    def work(t):
        try:
            for _ in range(5):
                out[t] = g.query_batch(qs[t], 10, nprobe=5 + t)
        except Exception as e:
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for t in range(4):
        _check("f32", out[t], refs[t], f"thread {t}")


@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_centroid_select_from_group_minima(gpu, metric):
    """The tensor-core centroid ranking selects from the minima of aligned groups of 8 cells (coarse_select_gm_kernel) and falls
    back to the full row of values when the selected groups hold more candidates than its buffer.  Two centroid tables, same
    data: cell ids in random order (the group path serves every query) and cell ids that follow a curve through space (a
    query's nearest cells are consecutive ids, every selected group is full of candidates, the full-row select takes over).
    Both must give the exact ranking's probe sets (src/cpu/ivf.rs:349-365), and the old full-row kernel (option
    `ivf_coarse_gm` = 0) the same again."""
    rng = np.random.default_rng(77)
    nlist, dim = 2048, 16
    t = np.linspace(0.0, 40.0, nlist, dtype=np.float64)
    curve = np.stack([np.cos(t * (1 + 0.1 * j)) * (1 + 0.05 * j) + 0.02 * t * (j % 3) for j in range(dim)], axis=1) + 3.0
    cent_curve = (curve + 1e-3 * rng.standard_normal((nlist, dim))).astype(np.float32)
    data = (cent_curve[rng.integers(0, nlist, 20000)] + 0.05 * rng.standard_normal((20000, dim))).astype(np.float32)
    q = (cent_curve[rng.integers(0, nlist, 300)] + 0.05 * rng.standard_normal((300, dim))).astype(np.float32)
    m_o = MET[metric][1]
    for name, cent in (("curve order", cent_curve), ("random order", np.ascontiguousarray(cent_curve[rng.permutation(nlist)]))):
        c = o.build_ivf(data, m_o, nlist=nlist, centroids=cent)
        g = _gpu_from_oracle(c)
        for k, nprobe in ((10, 8), (10, 32)):
            ref = o.ivf_search(c, q, k, nprobe=nprobe)
            for gm in (1, 0):
                g.set_option("ivf_coarse_gm", gm)
                got = g.query_batch(q, k, nprobe=nprobe)
                assert g.get_stat("coarse_path") == 2, (name, nprobe, gm)
                _check("f32", got, ref, f"{name} {metric} nprobe={nprobe} gm={gm}")
                assert g.get_stat("probed_lists") == int(ref[3].sum())


@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_device_lloyd_balanced_matches_restated_adjust_centers(gpu, metric):
    """KMeansTrainingParams::with_balancing: adjust_centers needs no random numbers, so from identical inputs
    the device loop and the oracle's restatement must move the same centroids: one balanced iteration is compared closely
    (same number of moves, centroids to 1e-6; the means are f64-accumulated on both sides in different orders, so the last bit
    of a mean may differ, which is also why LONG balanced runs are not compared step for step: one flipped borderline
    assignment changes which clusters starve).  The full run is checked for what balancing promises: fewer starved clusters
    than the plain loop."""
    rng = np.random.default_rng(11)
    e = np.eye(16, dtype=np.float32)      # one wide blob and two tiny ones in other directions (skewed under both metrics)
    data = np.concatenate([5 * e[0] + rng.normal(0, 0.5, (6000, 16)), 5 * e[1] + rng.normal(0, 0.02, (40, 16)),
                           -5 * e[1] + rng.normal(0, 0.02, (25, 16))]).astype(np.float32)
    label = np.concatenate([np.zeros(6000, int), np.ones(40, int), np.full(25, 2)])
    perm = rng.permutation(data.shape[0])
    data, label = np.ascontiguousarray(data[perm]), label[perm]
    init = np.ascontiguousarray(data[rng.choice(data.shape[0], 24, replace=False)])
    init[0], init[1] = data[np.nonzero(label == 1)[0][0]], data[np.nonzero(label == 2)[0][0]]   # two centroids that can only own 40 / 25 rows: starved
    m_g, m_o = (annb200.L2, o.L2) if metric == "l2" else (annb200.COSINE, o.COSINE)
    ref_c, ref_it, ref_moves = o.parallel_lloyd_balanced(data, init, m_o, 1, True, 42)
    c, it, moves = annb200.kmeans_lloyd(data, init, m_g, max_iters=1, balanced=True, seed=42, return_moves=True)
    assert (it, moves) == (ref_it, ref_moves), ((it, moves), (ref_it, ref_moves))
    assert moves > 0, "the skewed blobs must starve some clusters (otherwise balancing is not exercised)"
    assert np.allclose(c, ref_c, rtol=1e-6, atol=1e-6)
    # adjust_centers alone on identical inputs (assignments from the device, which equal the oracle's)
    cn = None if metric == "l2" else np.array([o.l2_norm_f32(v) for v in init], np.float32)
    assert np.array_equal(annb200.ivf_assign(data, init, m_g, cn), o.assign_all(data, init, cn, m_o))
    # full runs: balancing leaves fewer starved clusters than the plain loop, on the device as in the restatement
    def starved(cent):
        cnt = np.bincount(annb200.ivf_assign(data, cent, m_g), minlength=24)
        return int((cnt <= 0.25 * data.shape[0] / 24).sum())
    c_bal, it_bal, mv_bal = annb200.kmeans_lloyd(data, init, m_g, max_iters=30, balanced=True, seed=42, return_moves=True)
    c_plain, it_plain = annb200.kmeans_lloyd(data, init, m_g, max_iters=30)
    r_bal = o.parallel_lloyd_balanced(data, init, m_o, 30, True, 42)[0]
    assert mv_bal > 0 and np.isfinite(c_bal).all()
    assert starved(c_bal) <= starved(c_plain) and starved(r_bal) <= starved(c_plain)
    # balancing off through the same entry point = the plain loop, bit for bit
    c1, it1, mv1 = annb200.kmeans_lloyd(data, init, m_g, max_iters=30, balanced=False, return_moves=True)
    assert it_plain == it1 and mv1 == 0 and np.array_equal(c_plain.view(np.uint32), c1.view(np.uint32))


def test_kmeans_parallel_seeding_on_shared_draws(gpu):
    """k-means|| seeding with the D^2 passes on the device: on the same uniform draws the mirror must pick the
    same rows as a plain restatement of kmeans_parallel_init + weighted_kmeans_plus_plus built from the oracle's scalar
    kernels (src/utils/k_means_utils.rs:435-596)."""
    import math
    data = datagen.gaussian_noise(3000, 20, seed=17)
    k, seed = 12, 5
    got = annb200.kmeans_parallel_init(data, k, annb200.L2, np.random.Generator(np.random.PCG64(seed)))
    rng = np.random.Generator(np.random.PCG64(seed))
    n = data.shape[0]
    cand = [int(rng.integers(0, n))]
    for _ in range(int(math.log(k) + 1.0)):
        d = np.array([min(o.euclid_f32(v, data[c]) for c in cand) for v in data], dtype=np.float32)
        cs = np.cumsum(d.astype(np.float64))
        for _ in range(2 * k):
            cand.append(int(min(np.searchsorted(cs, float(rng.random()) * cs[-1], side="left"), n - 1)))
    cv = data[cand]
    chosen = [int(rng.integers(0, len(cand)))]
    dist = np.full(len(cand), np.inf, dtype=np.float32)
    for _ in range(1, k):
        dist = np.minimum(dist, np.array([o.euclid_f32(v, cv[chosen[-1]]) for v in cv], dtype=np.float32))
        cs = np.cumsum(dist.astype(np.float64))
        chosen.append(int(min(np.searchsorted(cs, float(rng.random()) * cs[-1], side="left"), len(cand) - 1)))
    assert np.array_equal(got.view(np.uint32), cv[chosen].view(np.uint32))
    # the build_* mirror with the reference's parameter struct as a dict
    ix = annb200.build_ivf_index_gpu(data, nlist=12, k_means_params={"iters": 5, "init": "kmeans||", "balanced": True}, dist_metric="euclidean", seed=3)
    ids, dist_, _ = ix.query_batch(data[:50], 5, nprobe=12)
    assert (ids[:, 0] == np.arange(50)).all()


@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_validate_index_matches_the_restated_trait(gpu, metric):
    """KnnValidation::validate_index (src/utils/mod.rs:210-242, src/cpu/ivf.rs:496-523) on shared sample positions: index
    query with the default nprobe against exhaustive_query over the stored (list-order) vectors, ids through original_ids."""
    data = datagen.gaussian_noise(30_000, 48, seed=51)
    c = o.build_ivf(data, MET[metric][1], nlist=200, kmeans_iters=4)
    g = _gpu_from_oracle(c)
    rng = np.random.default_rng(9)
    pos = rng.integers(0, c.n, size=300)
    k = 10
    got = g.validate_index(k, positions=pos)
    q = np.ascontiguousarray(c.vectors[pos])                     # stored rows, list order
    approx = o.ivf_search(c, q, k)[0]                             # default nprobe = sqrt(nlist)
    flat = o.build_flat(np.ascontiguousarray(c.vectors), MET[metric][1])
    true_pos = o.flat_search(flat, q, k)[0]
    true_ids = np.asarray(c.original_ids)[true_pos]
    want = np.mean([len(set(approx[i].tolist()) & set(true_ids[i].tolist())) / k for i in range(pos.size)])
    assert abs(got - want) < 1e-12, (got, want)
    assert 0.3 < got <= 1.0
    assert abs(g.validate_index(k, seed=3, no_samples=50) - g.validate_index(k, seed=3, no_samples=50)) == 0.0


@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("dim,k", [(160, 10), (256, 10), (200, 20)])
def test_wide_f32_rows_on_the_tensor_scan(gpu, metric, dim, k):
    """f32 lists with rows of 129 .. 256 elements on the tensor-core scan: hi query piece in TMEM, lo piece gathered into
    swizzled shared-memory slabs (SS-mode MMA for Qlo.Xhi)."""
    data = datagen.gaussian_noise(40_000, dim, seed=53)
    c = o.build_ivf(data, MET[metric][1], nlist=64, kmeans_iters=3)
    g = _gpu_from_oracle(c)
    g.set_option("path", annb200.PATH_TENSOR)
    q = datagen.subsample_with_noise(data, 500, seed=53)
    got = g.query_batch(q, k, nprobe=8)
    assert g.get_stat("last_path") == annb200.PATH_TENSOR
    ref = o.ivf_search(c, q, k, nprobe=8)
    _check("f32", got, ref[:3], f"ivf wide rows dim={dim} {metric}")


def test_centroid_tables_above_16384_cells(gpu):
    """nlist > 16 384: served by the ranked-prefix stages (tensor-core ranking + certified prefix, fused CUDA-core select); only a
    batch that needs the full per-row sort (a probe set outgrowing the prefix) is refused.  Centroids are supplied (random rows)."""
    rng = np.random.default_rng(61)
    data = datagen.gaussian_noise(120_000, 16, seed=61)
    nlist = 17_000
    cent = np.ascontiguousarray(data[rng.choice(data.shape[0], nlist, replace=False)])
    c = o.build_ivf(data, o.L2, nlist=nlist, centroids=cent)
    g = _gpu_from_oracle(c)
    q = datagen.subsample_with_noise(data, 300, seed=61)
    for nprobe in (24, 100):
        got = g.query_batch(q, 10, nprobe=nprobe)
        ref = o.ivf_search(c, q, 10, nprobe=nprobe)
        _check("f32", got, ref[:3], f"nlist {nlist} nprobe {nprobe}")


@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_ivf_f32_operand_forms_agree(gpu, metric):
    """The f32 tensor scan with 3xFP16 (default: pre-split fp16 hi / lo copy of the lists, rows scaled by powers of two) and with
    3xTF32 (in-kernel split of the raw rows) returns the oracle's rows; rows spanning many decades of norm."""
    rng = np.random.default_rng(73)
    base = datagen.correlated(40_000, 64, seed=73)
    data = np.ascontiguousarray(base * (np.float32(10.0) ** rng.integers(-4, 5, base.shape[0]).astype(np.float32))[:, None], dtype=np.float32)
    c = o.build_ivf(data, MET[metric][1], nlist=64, kmeans_iters=3)
    g = _gpu_from_oracle(c)
    g.set_option("path", annb200.PATH_TENSOR)
    q = datagen.subsample_with_noise(data, 600, seed=73)
    ref = o.ivf_search(c, q, 10, nprobe=8)
    assert g.get_stat("tc_kind") == 3
    bytes_fp16 = g.info().device_bytes
    _check("f32", g.query_batch(q, 10, nprobe=8), ref[:3], f"ivf 3xFP16 {metric}")
    assert g.get_stat("last_path") == annb200.PATH_TENSOR
    g.set_option("tc_f32_fp16", 0)
    assert g.get_stat("tc_kind") == 0 and g.info().device_bytes < bytes_fp16      # the fp16 copy of the lists is gone
    _check("f32", g.query_batch(q, 10, nprobe=8), ref[:3], f"ivf 3xTF32 {metric}")
    assert g.get_stat("last_path") == annb200.PATH_TENSOR
