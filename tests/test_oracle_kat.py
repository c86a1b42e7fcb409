"""Pins the CPU oracle (oracle/oracle.c) against the known-answer tests and golden
fixtures held by the reference's own unit tests.  Every test names the reference
test it restates (paths relative to /root/reference).  CPU only."""
import math

import numpy as np
import pytest

from oracle import oracle as o

SIMPLE = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1]], dtype=np.float32)  # create_simple_matrix


# ---------------------------------------------------------------- flat f32: src/cpu/exhaustive.rs:319-575
def test_flat_finds_self_euclidean_and_cosine():
    # test_exhaustive_query_finds_self_{euclidean,cosine} :345-372
    for metric in (o.L2, o.COSINE):
        ix = o.build_flat(SIMPLE, metric)
        ids, d, c = o.flat_search(ix, [[1, 0, 0]], 1)
        assert ids[0, 0] == 0 and abs(d[0, 0]) < 1e-5 and c[0] == 1


def test_flat_cosine_orthogonal_known_distances():
    # test_exhaustive_query_cosine_orthogonal :391-410
    ix = o.build_flat(SIMPLE, o.COSINE)
    ids, d, _ = o.flat_search(ix, [[1, 0, 0]], 5)
    want = [0.0, 1 - 1 / math.sqrt(2), 1 - 1 / math.sqrt(2), 1.0, 1.0]
    assert ids[0, 0] == 0
    np.testing.assert_allclose(d[0], want, atol=1e-5)
    assert set(ids[0, 1:3]) == {3, 4} and set(ids[0, 3:5]) == {1, 2}


def test_flat_k_larger_than_dataset():
    # test_exhaustive_query_k_larger_than_dataset :412-423 : exactly n results
    ix = o.build_flat(SIMPLE, o.L2)
    ids, d, c = o.flat_search(ix, [[1, 0, 0]], 10)
    assert c[0] == 5 and (ids[0, 5:] == -1).all() and np.isinf(d[0, 5:]).all()
    assert sorted(ids[0, :5].tolist()) == [0, 1, 2, 3, 4]


def test_flat_euclidean_known_distances():
    # test_exhaustive_euclidean_distances :437-459
    ix = o.build_flat(SIMPLE, o.L2)
    ids, d, _ = o.flat_search(ix, [[1, 0, 0]], 5)
    assert ids[0, 0] == 0
    np.testing.assert_allclose(d[0], [0, 1, 1, 2, 2], atol=1e-6)


def test_flat_all_points_found():
    # test_exhaustive_all_points_found :461-475
    ix = o.build_flat(SIMPLE, o.L2)
    ids, _, _ = o.flat_search(ix, [[0.5, 0.5, 0.5]], 5)
    assert sorted(ids[0].tolist()) == [0, 1, 2, 3, 4]


def test_flat_larger_dataset_formula_fixture():
    # test_exhaustive_larger_dataset :477-498 : data[i][j] = (i*j)/10, 50 x 10
    n, dim = 50, 10
    data = np.array([[(i * j) / 10.0 for j in range(dim)] for i in range(n)], dtype=np.float32)
    ix = o.build_flat(data, o.L2)
    ids, _, c = o.flat_search(ix, np.zeros((1, dim), np.float32), 5)
    assert c[0] == 5 and ids[0, 0] == 0


def test_flat_cosine_parallel_vectors():
    # test_exhaustive_cosine_parallel_vectors :500-524
    data = np.array([[1, 2, 3], [2, 4, 6], [-2, 1, 0]], dtype=np.float32)
    ix = o.build_flat(data, o.COSINE)
    ids, d, _ = o.flat_search(ix, [[1, 2, 3]], 3)
    assert ids[0].tolist() == [0, 1, 2]
    np.testing.assert_allclose(d[0], [0, 0, 1], atol=1e-5)


def test_flat_self_query_includes_self():
    # generate_knn (exhaustive.rs:255-292): self at rank 0 with distance 0
    ix = o.build_flat(SIMPLE, o.L2)
    ids, d, _ = o.flat_search(ix, None, 2, self_mode=True)
    assert (ids[:, 0] == np.arange(5)).all() and (d[:, 0] == 0).all()


def test_dimension_mismatch():
    # DimensionValidation (src/utils/traits.rs) -> AnnSearchErrors::DimensionMismatch
    ix = o.build_flat(SIMPLE, o.L2)
    with pytest.raises(ValueError):
        o.flat_search(ix, [[1, 0]], 1)


# ---------------------------------------------------------------- distance kernels: src/utils/dist.rs:5456-5720
def test_simd_kernels_match_scalar_within_1e5():
    rng = np.random.default_rng(0)
    for dim in (1, 3, 7, 8, 9, 16, 31, 32, 50, 100, 128, 129):
        a = rng.standard_normal(dim).astype(np.float32)
        b = rng.standard_normal(dim).astype(np.float32)
        assert abs(o.euclid_f32(a, b) - float(((a.astype(np.float64) - b) ** 2).sum())) < 1e-4
        assert abs(o.dot_f32(a, b) - float((a.astype(np.float64) * b).sum())) < 1e-4
        assert abs(o.l2_norm_f32(a) - float(np.sqrt((a.astype(np.float64) ** 2).sum()))) < 1e-4


def test_f32_lane_order_is_the_avx2_one():
    # euclidean_f32_avx2 (dist.rs:306-330): 8 lanes, mul + add, reduce_add tree, sequential tail
    rng = np.random.default_rng(1)
    a = rng.standard_normal(43).astype(np.float32)
    b = rng.standard_normal(43).astype(np.float32)
    acc = np.zeros(8, dtype=np.float32)
    for c in range(5):
        d = a[c * 8:(c + 1) * 8] - b[c * 8:(c + 1) * 8]
        acc = acc + d * d
    s = acc[:4] + acc[4:]
    tot = np.float32(np.float32(s[0] + s[2]) + np.float32(s[1] + s[3]))
    for e in range(40, 43):
        d = np.float32(a[e] - b[e])
        tot = np.float32(tot + np.float32(d * d))
    assert np.float32(o.euclid_f32(a, b)).view(np.uint32) == tot.view(np.uint32)


# ---------------------------------------------------------------- SQ8: src/utils/dist.rs:6212-6253, src/quantised/quantisers.rs:1001-1038
def test_sq8_euclidean_distance_i8():
    assert o.sq8_euclid(np.array([127, 0, 0], np.int8), np.array([127, 127, 0], np.int8)) == 16129.0


def test_sq8_cosine_distance_i8():
    q = np.array([127, 127, 0], np.int8)
    qn = 127 * 127 * 2
    assert abs(o.sq8_cosine(np.array([127, 0, 0], np.int8), 127 * 127, q, qn) - (1 - 1 / math.sqrt(2))) < 1e-5
    assert abs(o.sq8_cosine(np.array([127, 127, 0], np.int8), qn, q, qn)) < 1e-5
    assert o.sq8_cosine(np.array([0, 0, 0], np.int8), 0, q, qn) == 1.0   # zero norm -> 1.0 (dist.rs:5071-5075)


def test_scalar_quantiser_encode_decode():
    # test_scalar_quantiser_encode_decode :1001-1017
    sc = o.sq8_train(np.array([[127.0, 0.0, -127.0], [63.5, 0.0, -63.5]], np.float32))
    vec = np.array([[100.0, -25.0, 50.0]], np.float32)
    dec = o.sq8_decode(o.sq8_encode(vec, sc), sc)
    # dims 0 and 2 reconstruct within 2 %; dim 1 has the all-zero default scale 1.0
    assert abs(dec[0, 0] - 100.0) < 2.0 and abs(dec[0, 2] - 50.0) < 1.0 and abs(dec[0, 1] + 25.0) < 0.5 + 1e-6


def test_scalar_quantiser_clamping():
    # test_scalar_quantiser_clamping :1019-1029
    sc = o.sq8_train(np.array([[1.0, 1.0]], np.float32))
    enc = o.sq8_encode(np.array([[200.0, -200.0]], np.float32), sc)
    assert enc[0].tolist() == [127, -128]


def test_scalar_quantiser_zero_scale():
    # test_scalar_quantiser_zero_scale :1031-1038
    sc = o.sq8_train(np.array([[0.0, 10.0], [0.0, 20.0]], np.float32))
    assert sc[0] == 1.0 and abs(sc[1] - 20.0 / 128.0) < 1e-7


def test_sq8_rounding_is_half_away_from_zero_then_trunc():
    sc = np.array([1.0], np.float32)
    vals = np.array([[0.49], [0.5], [1.5], [-0.5], [-1.49], [0.0], [-0.0], [126.6], [-128.7]], np.float32)
    assert o.sq8_encode(vals, sc)[:, 0].tolist() == [0, 1, 2, -1, -1, 0, 0, 127, -128]


# ---------------------------------------------------------------- BF16: src/quantised/quantisers.rs:887-978
def test_bf16_encode_round_to_nearest_even():
    L = o.lib()
    assert L.orc_f32_to_bf16(1.0) == 0x3F80
    assert L.orc_f32_to_bf16(np.float32(1.00390625)) == 0x3F80   # exactly half way, even stays
    assert L.orc_f32_to_bf16(np.float32(1.01171875)) == 0x3F82   # half way, odd rounds up to even
    assert L.orc_f32_to_bf16(np.nextafter(np.float32(1.00390625), np.float32(2))) == 0x3F81   # just above half
    vals = np.array([1.0, -2.5, 3.7, 0.0, 100.5], np.float32)
    dec = o.decode_bf16(o.encode_bf16(vals))
    assert (np.abs(dec - vals) < 0.5).all() and (np.sign(dec) == np.sign(vals)).all()


def test_bf16_norm():
    # test_bf16_norm_vector :948-955
    assert abs(o.lib().orc_bf16_norm(o.encode_bf16(np.array([3.0, 4.0], np.float32)).ctypes.data_as(
        __import__("ctypes").c_void_p), 2) - 5.0) < 0.1


def test_bf16_flat_fixture():
    # src/quantised/exhaustive_bf16.rs:390-498 : data (i*dim+j)*0.1; self query finds self
    n, dim = 20, 8
    data = np.array([[(i * dim + j) * 0.1 for j in range(dim)] for i in range(n)], dtype=np.float32)
    ix = o.build_flat(data, o.L2, o.BF16)
    ids, d, _ = o.flat_search(ix, data[3:4], 3)
    assert ids[0, 0] == 3 and d[0, 0] < 0.05 and (np.diff(d[0]) >= 0).all()
    ids, d, _ = o.flat_search(ix, None, 2, self_mode=True)
    assert (ids[:, 0] == np.arange(n)).all() and (d[:, 0] == 0).all()


# ---------------------------------------------------------------- routing helpers: src/utils/k_means_utils.rs:3326-3480
def test_build_csr_layout():
    idx, off = o.build_csr([0, 1, 0, 2, 1, 0], 3)
    assert off.tolist() == [0, 3, 5, 6]
    assert idx[0:3].tolist() == [0, 2, 5] and idx[3:5].tolist() == [1, 4] and idx[5:6].tolist() == [3]
    _, off = o.build_csr([0, 0, 0], 1)
    assert off.tolist() == [0, 3]
    _, off = o.build_csr([0, 2, 0], 3)
    assert off.tolist() == [0, 2, 2, 3]


def test_select_probed_clusters_exact_orders():
    d = [0.5, 0.1, 0.9, 0.3]
    assert o.select_probed(d, [0, 1, 2, 3], [0, 5, 10, 15, 20], 3, 2) == [1, 3, 0]      # respects_nprobe_floor
    assert o.select_probed(d, [0, 1, 2, 3], [0, 2, 4, 6, 8], 1, 5) == [1, 3, 0]         # expands_to_reach_k
    assert o.select_probed([0.1, 0.5, 0.9], [0, 1, 2], [0, 0, 3, 6], 1, 2) == [0, 1]    # skips_empty_cells
    assert o.select_probed([0.5, 0.1, 0.9], [0, 1, 2], [0, 2, 4, 6], 1, 1000) == [1, 0, 2]  # caps_at_nlist


def test_assign_all_parallel():
    data = np.array([[0, 0], [0.1, 0.1], [10, 10], [9.9, 10.1]], np.float32)
    cent = np.array([[0, 0], [10, 10]], np.float32)
    assert o.assign_all(data, cent, np.ones(2, np.float32), o.L2).tolist() == [0, 0, 1, 1]
    data = np.array([[1, 0], [0, 1], [0.7, 0.1]], np.float32)
    cent = np.array([[1, 0], [0, 1]], np.float32)
    assert o.assign_all(data, cent, np.ones(2, np.float32), o.COSINE).tolist() == [0, 1, 0]


def test_assign_ties_pick_lowest_centroid():
    # strict `>` in direct_assign (k_means_utils.rs:2168-2171); flash_assign strict `<` (k_means_gpu.rs:177-376)
    cent = np.array([[1, 0], [1, 0], [0, 1]], np.float32)
    assert o.assign_all(np.array([[1, 0]], np.float32), cent, np.ones(3, np.float32), o.L2).tolist() == [0]


# ---------------------------------------------------------------- IVF: src/cpu/ivf.rs:549-835
def _singleton_ivf(metric=o.L2, dtype=o.F32):
    # one cluster per point (nlist = n = 5), centroids = the points themselves
    return o.build_ivf(SIMPLE, metric, nlist=5, dtype=dtype, centroids=SIMPLE.copy())


def test_ivf_query_returns_full_k_when_nprobe_underfills():
    # :777-803 probe expansion
    ix = _singleton_ivf()
    ids, d, c, npb, nsc = o.ivf_search(ix, [[1, 0, 0]], 3, nprobe=1)
    assert c[0] == 3 and ids[0, 0] == 0 and (np.diff(d[0]) >= 0).all()
    assert npb[0] == 3 and nsc[0] == 3


def test_ivf_finds_self_sorted_and_k_gt_n():
    ix = _singleton_ivf()
    ids, d, c, _, _ = o.ivf_search(ix, SIMPLE, 10, nprobe=5)
    assert (ids[:, 0] == np.arange(5)).all() and (c == 5).all()
    assert (np.diff(d[:, :5], axis=1) >= 0).all() and (ids[:, 5:] == -1).all()


def test_ivf_full_probe_equals_flat():
    # IVF with nprobe = nlist scans everything: same set as the exhaustive index (ivf.rs:734-754 fixture (i*j)/10)
    n, dim = 100, 10
    data = np.array([[(i * j) / 10.0 for j in range(dim)] for i in range(n)], dtype=np.float32)
    iv = o.build_ivf(data, o.L2, nlist=8)
    fl = o.build_flat(data, o.L2)
    q = data[::7] + np.float32(0.01)
    a = o.ivf_search(iv, q, 5, nprobe=8)
    b = o.flat_search(fl, q, 5)
    assert (np.sort(a[0], axis=1) == np.sort(b[0], axis=1)).all()
    assert (a[1].view(np.uint32) == b[1].view(np.uint32)).all()


def test_ivf_default_nprobe_and_self_scatter():
    rng = np.random.default_rng(3)
    data = rng.standard_normal((300, 16)).astype(np.float32)
    iv = o.build_ivf(data, o.COSINE, nlist=16)
    ids, d, c, npb, _ = o.ivf_search(iv, data[:4], 3)       # nprobe None -> floor(sqrt(16)) = 4
    assert (npb >= 4).all()
    ids, d, c, _, _ = o.ivf_search(iv, None, 3, nprobe=16, self_mode=True)
    assert (ids[:, 0] == np.arange(300)).all()              # generate_knn scatters to original ids (ivf.rs:476-486)


def test_ivf_sq8_and_bf16_run_and_find_self():
    rng = np.random.default_rng(4)
    data = rng.standard_normal((400, 24)).astype(np.float32) * 3
    for dtype in (o.BF16, o.SQ8):
        for metric in (o.L2, o.COSINE):
            iv = o.build_ivf(data, metric, nlist=10, dtype=dtype)
            ids, d, c, _, _ = o.ivf_search(iv, None, 4, nprobe=10, self_mode=True)
            assert (c == 4).all()
            assert (d[:, 0] <= d[:, 1]).all()
            # self is always among the zero-distance class at rank 0
            assert (np.abs(d[:, 0]) < 1e-2).all()


# ---------------------------------------------------------------- top-k order spec: src/gpu/topk_gpu.rs:1715-1731, 1931-2030
def _cpu_select(dists, k):
    order = sorted(range(len(dists)), key=lambda a: (dists[a], a))
    return order[:k]


@pytest.mark.parametrize("name,nq,mc,k,gen", [
    ("distinct", 8, 512, 50, lambda i: ((i * 7919 + 13) % 100003) * 0.001),
    ("all_duplicates", 4, 256, 30, lambda i: 0.25),
    ("heavy_ties", 4, 512, 40, lambda i: (i % 8) * 0.5),
])
def test_oracle_order_is_dist_then_position(name, nq, mc, k, gen):
    # A 1-d "database" whose squared distance to the query 0 reproduces the fixture values:
    # x = sqrt(d) so that (x - 0)^2 == d is not guaranteed bit-exact; use the values as coordinates of a
    # cosine-free L2 problem in which ties are exact: dist = (x)^2 with x taken from the fixture.
    for q in range(nq):
        vals = np.array([gen(q * mc + i) for i in range(mc)], dtype=np.float32)
        ix = o.build_flat(vals[:, None], o.L2)
        ids, d, _ = o.flat_search(ix, np.zeros((1, 1), np.float32), k)
        sq = (vals * vals).astype(np.float32)
        assert ids[0].tolist() == _cpu_select(sq.tolist(), k), name


# ---------------------------------------------------------------- pipeline fixture: src/gpu/dist_gpu.rs:1443-1583
@pytest.mark.parametrize("dim", [8, 32, 50])
def test_formula_pipeline_fixture(dim):
    nq, ndb, k = 10, 50, 5
    q = np.array([((i * 13 + 7) % 29) * 0.1 for i in range(nq * dim)], np.float32).reshape(nq, dim)
    db = np.array([((i * 17 + 3) % 31) * 0.1 for i in range(ndb * dim)], np.float32).reshape(ndb, dim)
    for metric in (o.L2, o.COSINE):
        ix = o.build_flat(db, metric)
        ids, d, _ = o.flat_search(ix, q, k)
        q64, db64 = q.astype(np.float64), db.astype(np.float64)
        if metric == o.L2:
            full = ((q64[:, None, :] - db64[None, :, :]) ** 2).sum(-1)
        else:
            full = 1 - (q64 @ db64.T) / (np.linalg.norm(q64, axis=1)[:, None] * np.linalg.norm(db64, axis=1)[None, :])
        want = np.sort(full, axis=1)[:, :k]
        np.testing.assert_allclose(d, want, rtol=1e-3, atol=1e-3)   # the reference's own tolerance (1e-3 / 1e-2)


def test_mega_kernel_known_answer():
    # src/gpu/dist_gpu.rs:1869-1947: 2 queries, 5 db vectors in two lists [0..3) and [3..5)
    db = np.array([[1, 1, 0, 0], [2, 0, 0, 0], [0, 0, 1, 1], [0, 2, 0, 0], [0, 0, 0, 3]], np.float32)
    q = np.array([[1, 0, 0, 0], [0, 1, 0, 0]], np.float32)
    ix = o.IvfIndex(o.F32, o.L2, 5, 4, 2, db, np.array([[1, 0.33, 0.33, 0.33], [0, 1, 0, 1.5]], np.float32),
                    np.array([0, 3, 5], np.int64), np.arange(5, dtype=np.int64))
    ids, d, c, _, _ = o.ivf_search(ix, q, 5, nprobe=2)
    assert d[0].tolist() == [1.0, 1.0, 3.0, 5.0, 10.0] and ids[0].tolist() == [0, 1, 2, 3, 4]
    ids, d, c, _, _ = o.ivf_search(ix, q[1:], 3, nprobe=1)
    assert c[0] == 3


def test_parallel_lloyd_restatement_stop_rule_and_means():
    """parallel_lloyd (k_means_utils.rs:1572-1700): convergence is tested before the update with a floor of
    max(1, n / 10000) changed assignments; a non-empty cluster's centroid is the mean of its members; an empty
    cluster keeps its centroid (:1657-1666)."""
    rng = np.random.default_rng(0)
    a = rng.normal(0.0, 0.1, (200, 4)).astype(np.float32)
    b = rng.normal(5.0, 0.1, (300, 4)).astype(np.float32)
    data = np.concatenate([a, b])
    init = np.stack([data[0], data[250], np.full(4, 100.0, np.float32)])
    cent, iters = o.parallel_lloyd(data, init, o.L2, max_iters=30)
    # iteration 0 assigns and updates, iteration 1 sees no change and stops before updating: one update
    assert iters == 1
    np.testing.assert_allclose(cent[0], a.astype(np.float64).mean(0), rtol=1e-6)
    np.testing.assert_allclose(cent[1], b.astype(np.float64).mean(0), rtol=1e-6)
    assert np.array_equal(cent[2], init[2])
    cent0, it0 = o.parallel_lloyd(data, init, o.L2, max_iters=0)
    assert it0 == 0 and np.array_equal(cent0, init)
