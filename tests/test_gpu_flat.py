"""GPU parity: flat (exhaustive) search through the C ABI against the CPU oracle.
Exact path (PATH_SIMT): bit-identical distances and identical ids for f32 / BF16 / SQ8, L2 / cosine,
external and self queries, ragged sizes, k > n, the reference's own fixtures."""
import numpy as np
import pytest

import annb200
from annb200 import datagen
from oracle import oracle as o
from util import assert_exact, assert_tolerance

pytestmark = pytest.mark.gpu

SIMPLE = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1]], dtype=np.float32)
DT = {"f32": (annb200.F32, o.F32), "bf16": (annb200.BF16, o.BF16), "sq8": (annb200.SQ8, o.SQ8)}
MET = {"l2": (annb200.L2, o.L2), "cosine": (annb200.COSINE, o.COSINE)}


def _pair(data, dtype, metric, path=annb200.PATH_SIMT):
    g = annb200.ExhaustiveIndexB200.new(data, MET[metric][0], DT[dtype][0])
    g.set_option("path", path)
    return g, o.build_flat(data, MET[metric][1], DT[dtype][1])


def test_simple_matrix_known_answers(gpu):
    # src/cpu/exhaustive.rs:319-533
    g, _ = _pair(SIMPLE, "f32", "cosine")
    ids, d, cnt = g.query_batch(np.array([[1, 0, 0]], np.float32), 5)
    assert ids[0, 0] == 0
    np.testing.assert_allclose(d[0], [0, 1 - 2 ** -0.5, 1 - 2 ** -0.5, 1, 1], atol=1e-5)
    g, _ = _pair(SIMPLE, "f32", "l2")
    ids, d, cnt = g.query_batch(np.array([[1, 0, 0]], np.float32), 10)   # k > n
    assert cnt[0] == 5 and (ids[0, 5:] == -1).all() and np.isinf(d[0, 5:]).all()
    np.testing.assert_allclose(d[0, :5], [0, 1, 1, 2, 2], atol=1e-6)


@pytest.mark.parametrize("dtype", ["f32", "bf16", "sq8"])
@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("n,dim,nq,k", [(5000, 32, 257, 15), (1237, 50, 33, 10), (3000, 128, 64, 10), (700, 7, 5, 31), (40, 3, 3, 64)])
def test_exact_parity_external_queries(gpu, dtype, metric, n, dim, nq, k):
    data = datagen.gaussian_noise(n, dim, seed=7)
    q = datagen.subsample_with_noise(data, nq, seed=7)
    g, c = _pair(data, dtype, metric)
    ids, d, cnt = g.query_batch(q, k)
    rids, rd, rcnt = o.flat_search(c, q, k)
    assert np.array_equal(cnt, rcnt)
    assert_exact(ids, d, rids, rd, f"flat {dtype} {metric} n={n} dim={dim}")
    assert g.get_stat("last_path") == annb200.PATH_SIMT


@pytest.mark.parametrize("dtype", ["f32", "bf16", "sq8"])
@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_exact_parity_self_queries(gpu, dtype, metric):
    data = datagen.correlated(2000, 50, seed=3)
    g, c = _pair(data, dtype, metric)
    ids, d, cnt = g.generate_knn(15)
    rids, rd, rcnt = o.flat_search(c, None, 15, self_mode=True)
    assert_exact(ids, d, rids, rd, f"flat self {dtype} {metric}")
    # sub-range of rows as the query set
    ids, d, cnt = g.generate_knn(5, row_begin=100, row_end=164)
    rids, rd, _ = o.flat_search(c, None, 5, self_rows=np.arange(100, 164), self_mode=True)
    assert_exact(ids, d, rids, rd, f"flat self sub-range {dtype} {metric}")


def test_config1_full_size_matches_oracle(gpu):
    """BASELINE config 1: flat f32 L2, 50k x 32 GaussianNoise, 1k queries, k = 15."""
    data = datagen.gaussian_noise(50_000, 32)
    q = datagen.subsample_with_noise(data, 1000)
    g, c = _pair(data, "f32", "l2")
    ids, d, _ = g.query_batch(q, 15)
    rids, rd, _ = o.flat_search(c, q, 15)
    assert_exact(ids, d, rids, rd, "config 1")
    assert_tolerance(ids, d, rids, rd, what="config 1 (tolerance form)")


def test_formula_fixture_and_ties(gpu):
    # src/gpu/dist_gpu.rs:1443-1583 data; src/gpu/topk_gpu.rs heavy-ties / all-duplicates order
    nq, ndb, dim, k = 10, 50, 8, 5
    q = np.array([((i * 13 + 7) % 29) * 0.1 for i in range(nq * dim)], np.float32).reshape(nq, dim)
    db = np.array([((i * 17 + 3) % 31) * 0.1 for i in range(ndb * dim)], np.float32).reshape(ndb, dim)
    for metric in ("l2", "cosine"):
        g, c = _pair(db, "f32", metric)
        assert_exact(*g.query_batch(q, k)[:2], *o.flat_search(c, q, k)[:2], f"formula {metric}")
    vals = np.array([(i % 8) * 0.5 for i in range(512)], np.float32)[:, None]      # heavy ties
    g, c = _pair(vals, "f32", "l2")
    assert_exact(*g.query_batch(np.zeros((1, 1), np.float32), 40)[:2], *o.flat_search(c, np.zeros((1, 1), np.float32), 40)[:2], "ties")
    same = np.full((256, 4), 0.5, np.float32)                                      # all duplicates: ids 0..k-1
    g, _ = _pair(same, "f32", "l2")
    ids, d, _ = g.query_batch(np.zeros((2, 4), np.float32), 30)
    assert (ids == np.arange(30)[None, :]).all()


def test_large_k_and_id_base(gpu):
    data = datagen.gaussian_noise(3000, 16, seed=11)
    q = datagen.subsample_with_noise(data, 9, seed=11)
    g = annb200.ExhaustiveIndexB200.new(data, annb200.L2, annb200.F32, id_base=1000)
    g.set_option("path", annb200.PATH_SIMT)
    c = o.build_flat(data, o.L2)
    for k in (1, 100, 250):
        ids, d, _ = g.query_batch(q, k)
        rids, rd, _ = o.flat_search(c, q, k)
        assert_exact(ids - 1000, d, rids, rd, f"k={k}")


def test_errors(gpu):
    g, _ = _pair(SIMPLE, "f32", "l2")
    with pytest.raises(annb200.AnnSearchError) as e:
        g.query_batch(np.zeros((2, 4), np.float32), 1)
    assert e.value.variant == "DimensionMismatch"
    with pytest.raises(annb200.AnnSearchError) as e:
        annb200.ExhaustiveIndexB200.new(SIMPLE, annb200.MANHATTAN)
    assert e.value.variant == "DistanceNotSupported"
    ids, d, cnt = g.query_batch(np.zeros((0, 3), np.float32), 3)     # empty query set
    assert ids.shape == (0, 3)


def test_determinism(gpu):
    data = datagen.gaussian_noise(4000, 32, seed=5)
    q = datagen.subsample_with_noise(data, 100, seed=5)
    g, _ = _pair(data, "sq8", "l2")
    a = g.query_batch(q, 20)
    b = g.query_batch(q, 20)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_mirror_free_functions(gpu):
    data = datagen.gaussian_noise(1000, 32, seed=9)
    q = datagen.subsample_with_noise(data, 20, seed=9)
    ix = annb200.build_exhaustive_index_gpu(data, "euclidean")
    ids, dist = annb200.query_exhaustive_index_gpu(q, ix, 5, return_dist=True)
    ids2, none = annb200.query_exhaustive_index_gpu(q, ix, 5, return_dist=False)
    assert none is None and np.array_equal(ids, ids2)
    ids, dist = annb200.query_exhaustive_index_gpu_self(ix, 3)
    assert (ids[:, 0] == np.arange(1000)).all()
    ram, vram = ix.memory_usage_bytes()
    assert vram >= 1000 * 32 * 4


def test_cpp_host_mirror_on_gpu(gpu):
    """host/annb200.hpp (C++ mirror of the reference API) through the C ABI: the reference's 5-point fixture."""
    import subprocess
    from test_abi import _build_cpp_mirror_test
    r = subprocess.run([_build_cpp_mirror_test()], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "OK gpu", r.stdout + r.stderr


def test_host_entry_batches_large_query_sets(gpu):
    """The host-buffer entry points walk the queries in internal batches of 16384: a 20 000-query call must equal the
    oracle row for row (flat tensor path and IVF list-major path)."""
    data = datagen.gaussian_noise(5000, 16, seed=91)
    q = datagen.subsample_with_noise(np.repeat(data, 4, axis=0), 20000, seed=91)
    g, c = _pair(data, "f32", "l2")
    ids, d, cnt = g.query_batch(q, 10)
    rids, rd, rcnt = o.flat_search(c, q, 10)
    assert_exact(ids, d, rids, rd, "flat, 20000 queries")
    ci = o.build_ivf(data, o.L2, nlist=64, kmeans_iters=4)
    gi = annb200.IvfIndexB200.from_parts(ci.vectors, ci.centroids, ci.offsets, ci.original_ids, ci.dtype, ci.metric)
    got = gi.query_batch(q, 10, nprobe=8)
    ref = o.ivf_search(ci, q, 10, nprobe=8)
    assert_exact(got[0], got[1], ref[0], ref[1], "ivf, 20000 queries")
    assert gi.get_stat("scanned_vectors") == int(ref[4].sum())


def test_concurrent_searches_on_one_handle(gpu):
    """`&self` queries are `Sync` in the reference (src/cpu/exhaustive.rs:142, src/gpu/exhaustive_gpu.rs:124): several host
    threads may search one index at once.  The handle serialises them internally; every thread gets the oracle's answer."""
    import threading
    data = datagen.gaussian_noise(20000, 32, seed=17)
    g, c = _pair(data, "f32", "l2", path=annb200.PATH_AUTO)
    qs = [datagen.subsample_with_noise(data, 300 + 50 * t, seed=100 + t) for t in range(4)]
    refs = [o.flat_search(c, q, 10) for q in qs]
    out, errs = [None] * 4, []

    def work(t):
        try:
            for _ in range(5):
                out[t] = g.query_batch(qs[t], 10)
        except Exception as e:   # surfaced below: an assertion inside a thread would be lost
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for t in range(4):
        assert_exact(out[t][0], out[t][1], refs[t][0], refs[t][1], f"thread {t}")


def test_matrix_to_flat_on_device(gpu):
    """annb_matrix_to_flat == matrix_to_flat (src/utils/mod.rs:44-68): any positively strided view -> row-major copy."""
    rng = np.random.default_rng(3)
    base = rng.standard_normal((1237, 70)).astype(np.float32)
    views = {
        "row-major": base,
        "column-major (faer default)": np.asfortranarray(base),
        "column-major, padded columns": np.asfortranarray(rng.standard_normal((1300, 70)).astype(np.float32))[:1237, :],
        "row-major, padded rows": base[:, :50],
        "both strides non-trivial": base[::2, ::3],
        "single column": np.asfortranarray(base)[:, :1],
        "single row": base[:1, :],
    }
    for name, v in views.items():
        got = annb200.matrix_to_flat(v)
        assert got.flags["C_CONTIGUOUS"] and np.array_equal(got.view(np.uint32), np.ascontiguousarray(v).view(np.uint32)), name
    with pytest.raises(annb200.AnnSearchError):
        annb200.matrix_to_flat(base[::-1])
    # the result feeds the index constructors: a column-major matrix and its row-major copy build the same index
    data = datagen.gaussian_noise(3000, 24, seed=5)
    q = datagen.subsample_with_noise(data, 40, seed=5)
    g = annb200.ExhaustiveIndexB200.new(annb200.matrix_to_flat(np.asfortranarray(data)), annb200.L2, annb200.F32)
    ids, d, _ = g.query_batch(q, 10)
    rids, rd, _ = o.flat_search(o.build_flat(data, o.L2, o.F32), q, 10)
    assert_exact(ids, d, rids, rd, "index built from a column-major matrix")


@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("shards", [1, 3])
def test_knn_graph_handoff(gpu, metric, shards):
    """the exact self-kNN graph in the shape of KnnGraphGpu (src/gpu/nndescent_gpu.rs:2418-2446): self edge
    dropped by id (duplicates of a row may precede it), rows ascending, sentinel pairs at the tail; the contract
    build_nsg_from_gpu_knn relies on holds; neighbours and distances are the oracle's self search without the self id."""
    data = datagen.gaussian_noise(6000, 24, seed=5)
    data[100] = data[40]                     # exact duplicates: row 40 precedes row 100 at distance 0
    data[5000] = data[40]
    k = 12
    dev = 0 if shards == 1 else [i % gpu for i in range(shards)]
    g = annb200.build_knn_graph_gpu(data, "cosine" if metric == "cosine" else "euclidean", k=k, device=dev)
    assert g.n == 6000 and g.dim == 24 and g.k == k and g.converged and g.vectors_flat.size == 6000 * 24
    assert (g.norms.size == 6000) == (metric == "cosine")
    assert g.check_contract()
    c = o.build_flat(data, o.COSINE if metric == "cosine" else o.L2)
    rids, rd, _ = o.flat_search(c, None, k + 1, self_mode=True)
    for i in range(6000):
        keep = rids[i] != i
        want_i, want_d = rids[i][keep][:k], rd[i][keep][:k]
        assert np.array_equal(g.pid[i], want_i) and np.array_equal(g.dist[i].view(np.uint32), want_d.view(np.uint32)), i
    # fewer than k other rows: sentinel padding
    small = annb200.build_knn_graph_gpu(data[:5], "euclidean", k=8)
    assert (small.pid[:, 4:] == annb200.SENTINEL_PID).all() and (small.dist[:, 4:] == np.finfo(np.float32).max).all() and small.check_contract()
