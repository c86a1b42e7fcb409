"""GPU parity of the tensor-core flat path (tcgen05 / TMEM / TMA pre-selection + exact re-rank).
The re-rank recomputes the k' survivors in the reference's own arithmetic, so results must be
bit-identical to the oracle whenever the pre-selection covers the true top-k."""
import numpy as np
import pytest

import annb200
from annb200 import datagen
from oracle import oracle as o
from util import assert_exact, assert_tolerance

pytestmark = pytest.mark.gpu

DT = {"f32": (annb200.F32, o.F32), "bf16": (annb200.BF16, o.BF16), "sq8": (annb200.SQ8, o.SQ8)}
MET = {"l2": (annb200.L2, o.L2), "cosine": (annb200.COSINE, o.COSINE)}


def _pair(data, dtype, metric):
    g = annb200.ExhaustiveIndexB200.new(data, MET[metric][0], DT[dtype][0])
    g.set_option("path", annb200.PATH_TENSOR)
    return g, o.build_flat(data, MET[metric][1], DT[dtype][1])


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("dim", [128, 32, 50])
def test_first_tile_values_match_a_float64_gemm(gpu, dtype, metric, dim):
    """The selection value of tile (0,0): v = |x|^2 - 2 q.x (L2) or -q.x/|x| (cosine)."""
    data = datagen.gaussian_noise(8192, dim, seed=3)
    q = datagen.subsample_with_noise(data, 128, seed=3)
    g, c = _pair(data, dtype, metric)
    g.set_option("tc_debug", 1)
    g.set_option("db_splits", 1)
    g.query_batch(q, 10)
    v = g.debug_fetch_tile().astype(np.float64)
    x = (o.decode_bf16(c.vectors[:128]) if dtype == "bf16" else data[:128]).astype(np.float64)
    s = q.astype(np.float64) @ x.T
    if metric == "l2":
        want = (x * x).sum(1)[None, :] - 2 * s
    else:
        want = -s / c.norms[:128].astype(np.float64)[None, :]
    scale = np.abs(want).max()
    err = np.abs(v - want).max() / scale
    assert err < 2e-6, f"tile error {err:.3e} (3xTF32 / bf16x3 should be ~1e-7)"


@pytest.mark.parametrize("dtype", ["f32", "bf16", "sq8"])
@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("n,dim,nq,k", [(20000, 128, 300, 10), (9000, 32, 129, 15), (5000, 50, 64, 10), (4100, 96, 7, 1)])
def test_exact_parity_with_oracle(gpu, dtype, metric, n, dim, nq, k):
    data = datagen.gaussian_noise(n, dim, seed=19)
    q = datagen.subsample_with_noise(data, nq, seed=19)
    g, c = _pair(data, dtype, metric)
    ids, d, cnt = g.query_batch(q, k)
    assert g.get_stat("last_path") == annb200.PATH_TENSOR
    rids, rd, rcnt = o.flat_search(c, q, k)
    assert_exact(ids, d, rids, rd, f"tensor flat {dtype} {metric} n={n} dim={dim} k={k}")
    for splits in (1, 3):
        g.set_option("db_splits", splits)
        ids, d, cnt = g.query_batch(q, k)
        assert_exact(ids, d, rids, rd, f"tensor flat db_splits={splits}")


@pytest.mark.parametrize("dtype", ["f32", "bf16", "sq8"])
def test_self_queries(gpu, dtype):
    data = datagen.correlated(6000, 64, seed=23)
    g, c = _pair(data, dtype, "cosine")
    ids, d, _ = g.generate_knn(10, row_begin=0, row_end=1000)
    rids, rd, _ = o.flat_search(c, None, 10, self_rows=np.arange(1000), self_mode=True)
    assert_exact(ids, d, rids, rd, f"tensor self {dtype}")


def test_correlated_cosine_hard_case(gpu):
    """BASELINE config 2 at reduced n: Correlated data, cosine -- nearest-neighbour gaps ~1e-6, the case 3xTF32 exists for."""
    data = datagen.correlated(100_000, 128, seed=42)
    q = datagen.subsample_with_noise(data, 1000, seed=42)
    g, c = _pair(data, "f32", "cosine")
    ids, d, _ = g.query_batch(q, 10)
    rids, rd, _ = o.flat_search(c, q, 10)
    assert_tolerance(ids, d, rids, rd, what="correlated cosine (tolerance contract)")
    assert_exact(ids, d, rids, rd, "correlated cosine")


def test_tensor_equals_simt_at_scale(gpu):
    """Size-independent property at a size the oracle would not finish quickly: both GPU paths agree bit for bit."""
    data = datagen.gaussian_noise(300_000, 128, seed=8)
    q = datagen.subsample_with_noise(data, 2000, seed=8)
    g = annb200.ExhaustiveIndexB200.new(data, annb200.L2, annb200.F32)
    g.set_option("path", annb200.PATH_TENSOR)
    a = g.query_batch(q, 10)
    g.set_option("path", annb200.PATH_SIMT)
    b = g.query_batch(q, 10)
    assert_exact(a[0], a[1], b[0], b[1], "tensor vs simt")


def test_certificate_and_exact_fallback(gpu):
    """With an absurdly pessimistic error bound every query fails the coverage certificate and is recomputed on the
    exact CUDA-core path; with the certificate off nothing is; results are identical either way."""
    data = datagen.gaussian_noise(12000, 64, seed=29)
    q = datagen.subsample_with_noise(data, 200, seed=29)
    g, c = _pair(data, "f32", "l2")
    ref = o.flat_search(c, q, 10)
    g.set_option("cert_eps_log2", -2)
    ids, d, _ = g.query_batch(q, 10)
    assert g.get_stat("uncertified") == 200 and g.get_stat("fallback_queries") == 200
    assert_exact(ids, d, ref[0], ref[1], "all queries through the fallback")
    g.set_option("cert_eps_log2", 0)
    ids, d, _ = g.query_batch(q, 10)
    assert g.get_stat("uncertified") == 0 and g.get_stat("fallback_queries") == 200
    assert_exact(ids, d, ref[0], ref[1], "certificate off")
    g.set_option("cert_eps_log2", -20)
    ids, d, _ = g.query_batch(q, 10)
    assert_exact(ids, d, ref[0], ref[1], "default bound")


@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("dim", [200, 400])
def test_sq8_int8_mma_wide_rows_and_ties(gpu, metric, dim):
    """SQ8 on the int8 tensor path (kind::i8, s32 accumulators): dims spanning several 128-code slabs, and coarse data whose
    integer distances tie heavily -- a tie between the k-th distance and the k'-th pre-selected value must be routed
    to the exact path, never resolved differently from the reference's (distance, id) order."""
    rng = np.random.default_rng(dim)
    data = rng.integers(-3, 4, size=(6000, dim)).astype(np.float32)          # few distinct codes -> many equal distances
    q = data[rng.choice(6000, 150, replace=False)] + rng.integers(-1, 2, size=(150, dim)).astype(np.float32)
    g, c = _pair(data, "sq8", metric)
    ref = o.flat_search(c, q, 10)
    for splits in (0, 1, 4):
        g.set_option("db_splits", splits)
        ids, d, _ = g.query_batch(q, 10)
        assert g.get_stat("last_path") == annb200.PATH_TENSOR
        assert_exact(ids, d, ref[0], ref[1], f"sq8 tensor {metric} dim={dim} splits={splits}")


def test_bf16_wide_rows_use_the_smem_query_operand(gpu):
    """bf16 rows wider than 128 do not fit the TMEM query budget (3 terms x 128 columns): the kernel falls back to
    shared-memory query operands and stays exact."""
    data = datagen.gaussian_noise(9000, 200, seed=31)
    q = datagen.subsample_with_noise(data, 140, seed=31)
    g, c = _pair(data, "bf16", "l2")
    ids, d, _ = g.query_batch(q, 10)
    assert g.get_stat("last_path") == annb200.PATH_TENSOR
    ref = o.flat_search(c, q, 10)
    assert_exact(ids, d, ref[0], ref[1], "bf16 dim 200")


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_second_chance_certificate(gpu, dtype):
    """Correlated data, L2: row norms dwarf the neighbour gaps, so with a conservative bound nearly every query fails the
    first coverage test (k'-th merged value).  The second test -- all candidates below the scan's final pruning
    threshold re-ranked, bound = that threshold -- must certify (almost) all of them, and results stay exact."""
    data = datagen.correlated(60_000, 128, seed=37)
    q = datagen.subsample_with_noise(data, 500, seed=37)
    g, c = _pair(data, dtype, "l2")
    ref = o.flat_search(c, q, 10)
    for e in (-18, -19):
        g.set_option("cert_eps_log2", e)
        ids, d, _ = g.query_batch(q, 10)
        assert g.get_stat("uncertified") <= 5, (e, g.get_stat("uncertified"))
        assert_exact(ids, d, ref[0], ref[1], f"second chance {dtype} eps=2^{e}")


@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("dim", [128, 50])
def test_bf16_hybrid_operand_placement(gpu, metric, dim):
    """Option tc_bf16_hybrid: the third bf16 term of the f32 query is multiplied from shared memory (SS-mode MMA) instead of
    TMEM.  Same selection values (tile (0,0) against a float64 product of the bf16-decoded rows) and the oracle's results."""
    data = datagen.gaussian_noise(20000, dim, seed=23)
    q = datagen.subsample_with_noise(data, 300, seed=23)
    g, c = _pair(data, "bf16", metric)
    g.set_option("tc_bf16_hybrid", 1)
    g.set_option("tc_debug", 1)
    g.set_option("db_splits", 1)
    ids, d, cnt = g.query_batch(q, 10)
    assert g.get_stat("last_path") == annb200.PATH_TENSOR
    v = g.debug_fetch_tile().astype(np.float64)
    x = o.decode_bf16(c.vectors[:128]).astype(np.float64)
    s = q[:128].astype(np.float64) @ x.T
    want = (x * x).sum(1)[None, :] - 2 * s if metric == "l2" else -s / c.norms[:128].astype(np.float64)[None, :]
    err = np.abs(v - want).max() / np.abs(want).max()
    assert err < 2e-6, f"tile error {err:.3e}"
    rids, rd, rcnt = o.flat_search(c, q, 10)
    assert_exact(ids, d, rids, rd, f"hybrid bf16 {metric} dim={dim}")
    g.set_option("tc_debug", 0)
    g.set_option("db_splits", 0)
    ids, d, cnt = g.query_batch(q, 10)
    assert_exact(ids, d, rids, rd, f"hybrid bf16 {metric} dim={dim}, auto splits")


def _adversarial_data(rng, n, dim, nq):
    """Worst case for the tensor core's accumulation and for the certificate: all-positive operands (every product and
    every partial sum has the same sign, so truncation errors add up instead of cancelling), dim 128, row norms spread
    over 10^3, and groups of near-duplicates whose mutual distances lie far BELOW the error bound (perturbation 1e-6) next
    to groups that are well separated (perturbation 3e-2)."""
    bases = rng.uniform(0.5, 1.0, size=(64, dim)).astype(np.float32)
    which = rng.integers(0, 64, n)
    scale = np.float32(10.0) ** rng.integers(0, 4, n).astype(np.float32)          # |x| in {1, 10, 100, 1000} x |base|
    pert = np.where(which < 32, np.float32(1e-6), np.float32(3e-2))
    rows = bases[which] * (1.0 + pert[:, None] * np.abs(rng.standard_normal((n, dim)))).astype(np.float32)
    data = np.ascontiguousarray(rows * scale[:, None], dtype=np.float32)
    pick = rng.choice(n, nq, replace=False)
    q = np.ascontiguousarray(data[pick] * (1.0 + np.float32(1e-3) * np.abs(rng.standard_normal((nq, dim)))).astype(np.float32))
    return data, q


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_adversarial_certificate_same_sign_operands(gpu, dtype, metric):
    """(i) With the exact fallback on, results are the oracle's bits -- whatever the certificate decided.
    (ii) With the fallback off, every query the certificate did NOT list must already be the oracle's bits:
    certified => equal.  (iii) The measured error of the selection values stays inside the bound the certificate assumes
    (at most one ulp of a |q||x|-sized partial sum per accumulating MMA)."""
    rng = np.random.default_rng(77)
    n, dim, nq, k = 16384, 128, 256, 10
    data, q = _adversarial_data(rng, n, dim, nq)
    g, c = _pair(data, dtype, metric)
    ref = o.flat_search(c, q, k)
    ids, d, _ = g.query_batch(q, k)
    assert g.get_stat("last_path") == annb200.PATH_TENSOR
    assert_exact(ids, d, ref[0], ref[1], f"adversarial {dtype} {metric}: with fallback")
    n_unc = g.get_stat("uncertified")
    g.set_option("cert_fallback", 0)
    ids0, d0, _ = g.query_batch(q, k)
    unc = set(g.uncertified_queries().tolist())
    assert len(unc) == n_unc
    differ = np.nonzero((ids0 != ref[0]).any(axis=1) | (bits(d0) != bits(ref[1])).any(axis=1))[0]
    assert set(differ.tolist()) <= unc, f"{len(set(differ.tolist()) - unc)} queries were certified but differ from the oracle"
    assert 0 < n_unc, "the near-duplicate groups must defeat the certificate (otherwise this test exercises nothing)"
    # (iii) tile (0, 0): selection values against float64 on the very operands the index stores
    g.set_option("cert_fallback", 1)
    g.set_option("tc_debug", 1)
    g.set_option("db_splits", 1)
    g.query_batch(q, k)
    v = g.debug_fetch_tile().astype(np.float64)
    x = (o.decode_bf16(c.vectors[:128]) if dtype == "bf16" else data[:128]).astype(np.float64)
    q64 = q[:128].astype(np.float64)
    s = q64 @ x.T
    qn, xn = np.sqrt((q64 * q64).sum(1)), np.sqrt((x * x).sum(1))
    eps = g.cert_eps()
    if metric == "l2":
        # f32 lists / rows go in as 3xFP16 pieces at ONE power-of-two scale for the whole index: their error is relative to the
        # largest row norm, the quantity the L2 bound eps * (|q| + |x|max)^2 is stated in; bf16 rows are stored exactly
        xs = np.full(128, np.sqrt((data.astype(np.float64) ** 2).sum(1)).max()) if dtype == "f32" else xn
        err = np.abs(v - ((x * x).sum(1)[None, :] - 2 * s)) / (qn[:, None] + xs[None, :]) ** 2
    else:
        err = np.abs(v - (-s / c.norms[:128].astype(np.float64)[None, :])) / qn[:, None]
    print(f"\n[adversarial {dtype} {metric}] max selection-value error {err.max():.3e} = {err.max() / eps:.3f} of the certificate bound {eps:.3e}; "
          f"uncertified {n_unc}/{nq}")
    assert err.max() <= eps, f"selection-value error {err.max():.3e} exceeds the bound {eps:.3e} the certificate assumes"


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ---- wide k (24 < k <= 128) and wide f32 rows (128 < dim <= 256) on the tensor path -------------------------------------
@pytest.mark.parametrize("dtype,metric,n,dim,nq,k", [
    ("f32", "l2", 40000, 64, 200, 32), ("f32", "cosine", 40000, 128, 150, 100), ("f32", "l2", 20000, 32, 130, 128),
    ("bf16", "l2", 30000, 96, 100, 64), ("sq8", "cosine", 30000, 64, 100, 50), ("f32", "l2", 5000, 50, 40, 25),
    ("f32", "cosine", 60000, 64, 100, 200), ("f32", "l2", 60000, 32, 70, 256)])
def test_wide_k_matches_oracle(gpu, dtype, metric, n, dim, nq, k):
    """k above one k' = 32 list: the union of the interleaved lists + the certificate against the smallest list threshold
    (reference spec of large k: src/gpu/topk_gpu.rs:992-1237, tested there at k = 250 / 1024; order (distance, index))."""
    data = datagen.gaussian_noise(n, dim, seed=31)
    q = datagen.subsample_with_noise(data, nq, seed=31)
    g, c = _pair(data, dtype, metric)
    ids, d, cnt = g.query_batch(q, k)
    assert g.get_stat("last_path") == annb200.PATH_TENSOR
    rids, rd, _ = o.flat_search(c, q, k)
    assert_exact(ids, d, rids, rd, f"wide-k tensor flat {dtype} {metric} n={n} dim={dim} k={k}")
    # randomly ordered rows spread a query's neighbours over all lists: (nearly) nothing may need the exact fallback
    assert g.get_stat("fallback_queries") <= nq // 20, g.get_stat("fallback_queries")


def test_wide_k_clustered_row_order_is_still_exact(gpu):
    """Rows sorted by cluster, clusters smaller than one list stride: a query's neighbours crowd into few lists, those lists'
    thresholds fall below the k-th distance, and the certificate must send such queries to the exact path instead of
    returning a wrong set."""
    rng = np.random.default_rng(5)
    centres = rng.normal(size=(400, 48)).astype(np.float32) * 6
    data = (np.repeat(centres, 40, axis=0) + rng.normal(size=(16000, 48)).astype(np.float32) * 0.05).astype(np.float32)   # 40 contiguous rows per cluster
    q = datagen.subsample_with_noise(data, 128, seed=5)
    g, c = _pair(data, "f32", "l2")
    ids, d, _ = g.query_batch(q, 60)
    rids, rd, _ = o.flat_search(c, q, 60)
    assert_exact(ids, d, rids, rd, "wide-k, clustered row order")


@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("dim,k", [(160, 10), (256, 15), (200, 40), (132, 1)])
def test_wide_f32_rows_on_the_tensor_path(gpu, metric, dim, k):
    """f32 rows of more than 128 elements: hi query piece in TMEM, lo piece in shared memory (SS-mode MMA for Qlo.Xhi).
    The reference accepts any dim (src/gpu/exhaustive_gpu.rs:116-117)."""
    data = datagen.gaussian_noise(12000, dim, seed=37)
    q = datagen.subsample_with_noise(data, 140, seed=37)
    g, c = _pair(data, "f32", metric)
    ids, d, _ = g.query_batch(q, k)
    assert g.get_stat("last_path") == annb200.PATH_TENSOR
    rids, rd, _ = o.flat_search(c, q, k)
    assert_exact(ids, d, rids, rd, f"wide f32 rows dim={dim} {metric} k={k}")


def test_wide_f32_rows_first_tile_values(gpu):
    data = datagen.gaussian_noise(8192, 256, seed=3)
    q = datagen.subsample_with_noise(data, 128, seed=3)
    g, c = _pair(data, "f32", "l2")
    g.set_option("tc_debug", 1)
    g.set_option("db_splits", 1)
    g.query_batch(q, 10)
    v = g.debug_fetch_tile().astype(np.float64)
    x = data[:128].astype(np.float64)
    want = (x * x).sum(1)[None, :] - 2 * (q.astype(np.float64) @ x.T)
    err = np.abs(v - want).max() / np.abs(want).max()
    # the tensor core's f32 accumulation truncates: the error grows with the MMAs per tile row (96 here: 3.7e-6 measured,
    # 1.6e-6 at dim 128, 1.0e-6 at dim 64 -- tools/tile_err.py; exact 3xTF32 arithmetic would give 3e-8); tc_cert_eps budgets it
    assert err < 6e-6, f"tile error {err:.3e}"


@pytest.mark.parametrize("opt", ["tc_f32_lo_smem", "tc_strided"])
def test_layout_options_do_not_change_results(gpu, opt):
    data = datagen.correlated(30000, 128, seed=41)
    q = datagen.subsample_with_noise(data, 300, seed=41)
    g, c = _pair(data, "f32", "cosine")
    ref = o.flat_search(c, q, 10)
    g.set_option(opt, 1)
    ids, d, _ = g.query_batch(q, 10)
    assert g.get_stat("last_path") == annb200.PATH_TENSOR
    assert_exact(ids, d, ref[0], ref[1], opt)


@pytest.mark.parametrize("dtype,metric,dim,k", [("f32", "l2", 320, 10), ("f32", "cosine", 512, 15), ("f32", "l2", 388, 50),
                                                ("bf16", "cosine", 384, 10), ("bf16", "l2", 1000, 10), ("sq8", "l2", 640, 10), ("sq8", "cosine", 2048, 12)])
def test_streamed_query_slabs_for_very_wide_rows(gpu, dtype, metric, dim, k):
    """Rows too wide for a resident query tile (f32 dim > 256, bf16 > 256, int8 > 512; up to 2048 B per row): the query slabs
    stream through the ring with the database slabs (SS-mode MMAs)."""
    data = datagen.gaussian_noise(9000, dim, seed=43)
    q = datagen.subsample_with_noise(data, 130, seed=43)
    g, c = _pair(data, dtype, metric)
    ids, d, _ = g.query_batch(q, k)
    assert g.get_stat("last_path") == annb200.PATH_TENSOR
    rids, rd, _ = o.flat_search(c, q, k)
    assert_exact(ids, d, rids, rd, f"streamed query slabs {dtype} dim={dim} {metric} k={k}")


def test_escalation_to_wide_mode_after_an_uncertified_batch(gpu):
    """A handle whose batch left > 2 % of the queries uncertified runs its later batches in wide-k mode (stat tc_escalated);
    forced here with a pessimistic error bound.  Results stay the oracle's in both modes."""
    data = datagen.gaussian_noise(20000, 96, seed=47)
    q = datagen.subsample_with_noise(data, 400, seed=47)
    g, c = _pair(data, "f32", "l2")
    ref = o.flat_search(c, q, 10)
    g.set_option("cert_eps_log2", -6)
    for level in (1, 2):     # k' = 32, then wide-k mode
        ids, d, _ = g.query_batch(q, 10)
        assert g.get_stat("tc_escalated") == level and g.get_stat("fallback_queries") > 8
        assert_exact(ids, d, ref[0], ref[1], f"escalation level {level}")
    g.set_option("cert_eps_log2", 1)
    ids, d, _ = g.query_batch(q, 10)
    assert g.get_stat("last_path") == annb200.PATH_TENSOR
    assert_exact(ids, d, ref[0], ref[1], "wide mode at k = 10")


@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_f32_operand_forms_agree(gpu, metric):
    """3xFP16 (default: fp16 hi + lo of operands scaled by powers of two) and 3xTF32 pre-selection give the oracle's rows; data
    with a wide dynamic range inside rows and across rows.  The selection values must stay inside the certificate's error model:
    cosine operands are scaled row by row (unit rows), so their error is relative to |q| alone; L2 database operands carry ONE scale
    for the whole index, so their error is relative to (|q| + |x|max)^2 with the LARGEST row norm -- the quantity the L2 bound is
    stated in -- and rows far below it are exact only through the re-rank / fallback, which assert_exact checks."""
    rng = np.random.default_rng(71)
    base = datagen.correlated(30_000, 96, seed=71)
    row_scale = np.float32(10.0) ** rng.integers(-6, 7, base.shape[0]).astype(np.float32)            # |x| over 12 decades
    col_scale = np.float32(2.0) ** rng.integers(-10, 11, base.shape[1]).astype(np.float32)           # elements over 6 decades inside a row
    data = np.ascontiguousarray(base * row_scale[:, None] * col_scale[None, :], dtype=np.float32)
    q = datagen.subsample_with_noise(data, 300, seed=71)
    g, c = _pair(data, "f32", metric)
    ref = o.flat_search(c, q, 10)
    assert g.get_stat("tc_kind") == 3
    ids, d, _ = g.query_batch(q, 10)
    assert g.get_stat("last_path") == annb200.PATH_TENSOR
    assert_exact(ids, d, ref[0], ref[1], f"3xFP16 {metric}")
    g.set_option("tc_f32_fp16", 0)
    assert g.get_stat("tc_kind") == 0
    ids, d, _ = g.query_batch(q, 10)
    assert_exact(ids, d, ref[0], ref[1], f"3xTF32 {metric}")
    g.set_option("tc_f32_fp16", 1)
    g.set_option("tc_debug", 1)
    g.set_option("db_splits", 1)
    g.query_batch(q, 10)
    v = g.debug_fetch_tile().astype(np.float64)
    x = data[:128].astype(np.float64)
    s = q[:128].astype(np.float64) @ x.T
    qn, xn = np.sqrt((q[:128].astype(np.float64) ** 2).sum(1)), np.sqrt((x * x).sum(1))
    if metric == "l2":
        xn_max = np.sqrt((data.astype(np.float64) ** 2).sum(1)).max()
        err = np.abs(v - ((x * x).sum(1)[None, :] - 2 * s)) / (qn[:, None] + xn_max) ** 2
    else:
        err = np.abs(v - (-s / c.norms[:128].astype(np.float64)[None, :])) / qn[:, None]
    eps = g.cert_eps()
    assert err.max() <= eps, f"selection-value error {err.max():.3e} exceeds the certificate bound {eps:.3e}"
