"""ctypes front-end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product (ann-search-rs_b200/, include/) never imports this module.

The index-building helpers below follow the reference constructors step by step
(file:line cited per function) with two stated stand-ins: numpy's PCG64 replaces
rand's StdRng for the training subsample, and `orc_kmeans_lloyd` replaces
`train_centroids` (neither stream is reproducible without a Rust toolchain).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

L2, COSINE = 0, 1
F32, BF16, SQ8 = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so (gcc, a few seconds)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "liboracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        f = C.c_float
        _lib.orc_euclid_f32.restype = f
        _lib.orc_dot_f32.restype = f
        _lib.orc_l2_norm_f32.restype = f
        _lib.orc_seq_norm_f32.restype = f
        _lib.orc_euclid_bf16_f32.restype = f
        _lib.orc_dot_bf16_f32.restype = f
        _lib.orc_euclid_bf16_bf16.restype = f
        _lib.orc_dot_bf16_bf16.restype = f
        _lib.orc_bf16_norm.restype = f
        _lib.orc_bf16_to_f32.restype = f
        _lib.orc_bf16_to_f32.argtypes = [C.c_uint16]
        _lib.orc_f32_to_bf16.restype = C.c_uint16
        _lib.orc_f32_to_bf16.argtypes = [C.c_float]
        _lib.orc_sq8_euclid.restype = f
        _lib.orc_sq8_cosine.restype = f
        _lib.orc_sq8_norm_sq.restype = C.c_int32
        _lib.orc_select_probed.restype = C.c_int
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _p(a, ty=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def max_threads() -> int:
    return int(lib().orc_max_threads())


# --------------------------------------------------------------------------
# scalar kernels
# --------------------------------------------------------------------------
def euclid_f32(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_euclid_f32(_p(a), _p(b), C.c_int(a.size)))


def dot_f32(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_dot_f32(_p(a), _p(b), C.c_int(a.size)))


def l2_norm_f32(a):
    a = _f32(a)
    return float(lib().orc_l2_norm_f32(_p(a), C.c_int(a.size)))


def row_norms_f32(x):
    x = _f32(x)
    out = np.empty(x.shape[0], dtype=np.float32)
    lib().orc_row_norms_f32(_p(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]), _p(out))
    return out


def seq_norm_f32(a):
    a = _f32(a)
    return float(lib().orc_seq_norm_f32(_p(a), C.c_int(a.size)))


def normalise_rows(x):
    x = _f32(x).copy()
    lib().orc_normalise_rows_f32(_p(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]))
    return x


def encode_bf16(x):
    x = _f32(x)
    out = np.empty(x.shape, dtype=np.uint16)
    lib().orc_encode_bf16(_p(x), _p(out), C.c_int64(x.size))
    return out


def decode_bf16(b):
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def sq8_train(x):
    x = _f32(x)
    scales = np.empty(x.shape[1], dtype=np.float32)
    lib().orc_sq8_train(_p(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]), _p(scales))
    return scales


def sq8_encode(x, scales):
    x = _f32(np.atleast_2d(x))
    scales = _f32(scales)
    out = np.empty(x.shape, dtype=np.int8)
    lib().orc_sq8_encode_all(_p(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]), _p(scales), _p(out))
    return out


def sq8_decode(codes, scales):
    return codes.astype(np.float32) * _f32(scales)[None, :]


def sq8_euclid(db, q):
    db = np.ascontiguousarray(db, dtype=np.int8)
    q = np.ascontiguousarray(q, dtype=np.int8)
    return float(lib().orc_sq8_euclid(_p(db), _p(q), C.c_int(q.size)))


def sq8_cosine(db, db_norm_sq, q, q_norm_sq):
    db = np.ascontiguousarray(db, dtype=np.int8)
    q = np.ascontiguousarray(q, dtype=np.int8)
    return float(lib().orc_sq8_cosine(_p(db), C.c_int32(db_norm_sq), _p(q), C.c_int32(q_norm_sq), C.c_int(q.size)))


def build_csr(assign, nlist):
    assign = np.ascontiguousarray(assign, dtype=np.int64)
    n = assign.size
    idx = np.empty(n, dtype=np.int64)
    off = np.empty(nlist + 1, dtype=np.int64)
    lib().orc_build_csr(_p(assign), C.c_int64(n), C.c_int(nlist), _p(idx), _p(off))
    return idx, off


def select_probed(dists, cells, offsets, nprobe, k):
    dists = _f32(dists)
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    out = np.empty(dists.size, dtype=np.int32)
    cnt = lib().orc_select_probed(_p(dists), _p(cells), C.c_int(dists.size), _p(offsets), C.c_int(nprobe),
                                  C.c_int64(k), _p(out))
    return out[:cnt].tolist()


def assign_all(data, centroids, centroid_norms, metric, nthreads=0):
    data, centroids = _f32(data), _f32(centroids)
    nlist = centroids.shape[0]
    cn = _f32(centroid_norms) if centroid_norms is not None else np.ones(nlist, dtype=np.float32)
    out = np.empty(data.shape[0], dtype=np.int64)
    lib().orc_assign_all(_p(data), C.c_int64(data.shape[0]), C.c_int(data.shape[1]), _p(centroids), _p(cn),
                         C.c_int(nlist), C.c_int(metric), _p(out), C.c_int(nthreads))
    return out


def kmeans_lloyd(train, nlist, metric, iters=10, nthreads=0):
    train = _f32(train)
    cent = np.empty((nlist, train.shape[1]), dtype=np.float32)
    lib().orc_kmeans_lloyd(_p(train), C.c_int64(train.shape[0]), C.c_int(train.shape[1]), C.c_int(nlist),
                           C.c_int(metric), C.c_int(iters), _p(cent), C.c_int(nthreads))
    return cent


def parallel_lloyd(data, init_centroids, metric, max_iters=30, nthreads=0):
    """Unbalanced parallel_lloyd (src/utils/k_means_utils.rs:1572-1700) from given initial centroids.
    Returns (centroids, number of updates performed)."""
    data = _f32(data)
    cent = _f32(init_centroids).copy()
    L = lib()
    L.orc_parallel_lloyd.restype = C.c_int
    it = L.orc_parallel_lloyd(_p(data), C.c_int64(data.shape[0]), C.c_int(data.shape[1]), C.c_int(cent.shape[0]), C.c_int(metric),
                              C.c_int(max_iters), _p(cent), C.c_int(nthreads))
    return cent, int(it)


def adjust_centers(centroids, data, assign, counts, seed):
    """adjust_centers (src/utils/k_means_utils.rs:979-1030).  Returns (moved centroids copy, number moved)."""
    cent = _f32(centroids).copy()
    data = _f32(data)
    a = np.ascontiguousarray(assign, dtype=np.int64)
    c = np.ascontiguousarray(counts, dtype=np.int64)
    L = lib()
    L.orc_adjust_centers.restype = C.c_int64
    moved = L.orc_adjust_centers(_p(cent), C.c_int(cent.shape[1]), C.c_int(cent.shape[0]), _p(data), C.c_int64(data.shape[0]), _p(a), _p(c),
                                 C.c_uint64(seed))
    return cent, int(moved)


def parallel_lloyd_balanced(data, init_centroids, metric, max_iters=30, balanced=True, seed=42, nthreads=0):
    """parallel_lloyd with the balancing hook (src/utils/k_means_utils.rs:1572-1700).  Returns (centroids, updates, centroid moves)."""
    data = _f32(data)
    cent = _f32(init_centroids).copy()
    L = lib()
    L.orc_parallel_lloyd_balanced.restype = C.c_int
    adj = C.c_int64(0)
    it = L.orc_parallel_lloyd_balanced(_p(data), C.c_int64(data.shape[0]), C.c_int(data.shape[1]), C.c_int(cent.shape[0]), C.c_int(metric),
                                       C.c_int(max_iters), C.c_int(1 if balanced else 0), C.c_uint64(seed), _p(cent), C.byref(adj), C.c_int(nthreads))
    return cent, int(it), int(adj.value)


# --------------------------------------------------------------------------
# index containers (host mirrors of the reference structs)
# --------------------------------------------------------------------------
@dataclass
class FlatIndex:
    """ExhaustiveIndex / ExhaustiveIndexBf16 / ExhaustiveSq8Index
    (src/cpu/exhaustive.rs:83-105, src/quantised/exhaustive_bf16.rs:94-121,
    src/quantised/exhaustive_sq8.rs:104-151)."""
    dtype: int
    metric: int
    n: int
    dim: int
    vectors: np.ndarray                 # f32 / uint16 / int8, row-major
    norms: Optional[np.ndarray] = None  # f32 (cosine, f32+bf16)
    norms_i: Optional[np.ndarray] = None  # int32 (cosine, sq8)
    scales: Optional[np.ndarray] = None   # f32[dim] (sq8)


def build_flat(data, metric, dtype=F32) -> FlatIndex:
    data = _f32(data)
    n, dim = data.shape
    if dtype in (F32, BF16):
        norms = row_norms_f32(data) if metric == COSINE else None
        vec = data if dtype == F32 else encode_bf16(data)
        return FlatIndex(dtype, metric, n, dim, vec, norms)
    x = normalise_rows(data) if metric == COSINE else data
    scales = sq8_train(x)
    codes = sq8_encode(x, scales)
    norms_i = (codes.astype(np.int32) ** 2).sum(axis=1).astype(np.int32) if metric == COSINE else None
    return FlatIndex(SQ8, metric, n, dim, codes, None, norms_i, scales)


def flat_search(ix: FlatIndex, queries, k, self_rows=None, self_mode=False, nthreads=0, return_dist=True):
    L = lib()
    if self_mode:
        nq = ix.n if self_rows is None else len(self_rows)
        q = None
        sr = None if self_rows is None else np.ascontiguousarray(self_rows, dtype=np.int64)
    else:
        q = _f32(queries)
        if q.ndim != 2 or q.shape[1] != ix.dim:
            raise ValueError("DimensionMismatch")  # src/utils/traits.rs DimensionValidation
        nq = q.shape[0]
        sr = None
    ids = np.empty((nq, k), dtype=np.int64)
    dist = np.empty((nq, k), dtype=np.float32)
    cnt = np.empty(nq, dtype=np.int32)
    L.orc_flat_search(C.c_int(ix.dtype), C.c_int(ix.metric), _p(ix.vectors), C.c_int64(ix.n), C.c_int(ix.dim),
                      _p(ix.norms), _p(ix.norms_i), _p(ix.scales), _p(q), C.c_int64(nq), _p(sr),
                      C.c_int(1 if self_mode else 0), C.c_int(k), _p(ids), _p(dist), _p(cnt), C.c_int(nthreads))
    return ids, dist, cnt


@dataclass
class IvfIndex:
    """IvfIndex / IvfIndexBf16 / IvfSq8Index after optimise_memory_layout
    (src/cpu/ivf.rs:25-48, 257-294)."""
    dtype: int
    metric: int
    n: int
    dim: int
    nlist: int
    vectors: np.ndarray                  # list order
    centroids: np.ndarray                # f32 [nlist, dim]
    offsets: np.ndarray                  # int64 [nlist+1]
    original_ids: np.ndarray             # int64 [n], new -> old
    norms: Optional[np.ndarray] = None
    centroid_norms: Optional[np.ndarray] = None
    norms_i: Optional[np.ndarray] = None
    scales: Optional[np.ndarray] = None
    extra: dict = field(default_factory=dict)


def build_ivf(data, metric, nlist=None, dtype=F32, seed=42, kmeans_iters=10, centroids=None,
              nthreads=0) -> IvfIndex:
    """IvfIndex::build (src/cpu/ivf.rs:145-249), IvfIndexBf16::build
    (src/quantised/ivf_bf16.rs:150-254), IvfSq8Index::build
    (src/quantised/ivf_sq8.rs:158-284)."""
    data = _f32(data)
    n, dim = data.shape
    if nlist is None:
        nlist = int(np.float32(n) ** np.float32(0.5))
    nlist = max(int(nlist), 1)
    x = data
    if dtype == SQ8 and metric == COSINE:
        x = normalise_rows(data)                       # ivf_sq8.rs:169-176
    norms = row_norms_f32(data) if (metric == COSINE and dtype != SQ8) else None
    n_train = max(min(256 * nlist, 250_000, n), 1)      # ivf.rs:174
    rng = np.random.Generator(np.random.PCG64(seed))
    perm = rng.permutation(n)[:n_train]                 # stand-in for sample_vectors (k_means_utils.rs:3047-3069)
    train = np.ascontiguousarray(x[perm])
    if centroids is None:
        centroids = kmeans_lloyd(train, nlist, metric, kmeans_iters, nthreads)
    centroids = _f32(centroids).copy()
    scales = None
    if dtype == SQ8:
        if metric == COSINE:
            centroids = normalise_rows(centroids)       # ivf_sq8.rs:198-205
        scales = sq8_train(train)                       # codebook from the training sample only (:211)
        cn = np.ones(nlist, dtype=np.float32)
        assign = assign_all(x, centroids, cn, metric, nthreads)
        centroid_norms = None
    else:
        centroid_norms = (np.array([seq_norm_f32(c) for c in centroids], dtype=np.float32)
                          if metric == COSINE else None)  # ivf.rs:193-206 (sequential fold)
        assign = assign_all(x, centroids, centroid_norms, metric, nthreads)
    all_idx, offsets = build_csr(assign, nlist)
    new_to_old = all_idx                                 # optimise_memory_layout: lists in order (ivf.rs:257-294)
    if dtype == F32:
        vec = np.ascontiguousarray(x[new_to_old])
    elif dtype == BF16:
        vec = np.ascontiguousarray(encode_bf16(x)[new_to_old])
    else:
        vec = np.ascontiguousarray(sq8_encode(x, scales)[new_to_old])
    norms_l = np.ascontiguousarray(norms[new_to_old]) if norms is not None else None
    norms_i = None
    if dtype == SQ8 and metric == COSINE:
        norms_i = (vec.astype(np.int32) ** 2).sum(axis=1).astype(np.int32)
    return IvfIndex(dtype, metric, n, dim, nlist, vec, centroids, offsets, new_to_old.astype(np.int64),
                    norms_l, centroid_norms, norms_i, scales)


def ivf_search(ix: IvfIndex, queries, k, nprobe=None, self_rows=None, self_mode=False, nthreads=0):
    L = lib()
    if self_mode:
        nq = ix.n if self_rows is None else len(self_rows)
        q = None
        sr = None if self_rows is None else np.ascontiguousarray(self_rows, dtype=np.int64)
    else:
        q = _f32(queries)
        if q.ndim != 2 or q.shape[1] != ix.dim:
            raise ValueError("DimensionMismatch")
        nq = q.shape[0]
        sr = None
    ids = np.empty((nq, k), dtype=np.int64)
    dist = np.empty((nq, k), dtype=np.float32)
    cnt = np.empty(nq, dtype=np.int32)
    npb = np.empty(nq, dtype=np.int32)
    nsc = np.empty(nq, dtype=np.int64)
    L.orc_ivf_search(C.c_int(ix.dtype), C.c_int(ix.metric), _p(ix.vectors), C.c_int64(ix.n), C.c_int(ix.dim),
                     _p(ix.norms), _p(ix.norms_i), _p(ix.scales), _p(ix.centroids), _p(ix.centroid_norms),
                     C.c_int(ix.nlist), _p(ix.offsets), _p(ix.original_ids), _p(q), C.c_int64(nq), _p(sr),
                     C.c_int(1 if self_mode else 0), C.c_int(k), C.c_int(nprobe if nprobe else 0),
                     _p(ids), _p(dist), _p(cnt), _p(npb), _p(nsc), C.c_int(nthreads))
    return ids, dist, cnt, npb, nsc


def recall_at_k(true_ids, approx_ids, k):
    """examples/commons/mod.rs:923-940 calculate_recall."""
    tot = 0.0
    for t, a in zip(true_ids, approx_ids):
        ts = set(int(v) for v in t[:k])
        as_ = set(int(v) for v in a[:k])
        tot += len(ts & as_) / float(k)
    return tot / len(true_ids)
