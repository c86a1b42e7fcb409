"""annb200 -- ctypes binding of libannb200 and a host-side mirror of the reference API.

The function names, argument meaning and error behaviour follow the free functions
of ann-search-rs' src/lib.rs for the flat and IVF families:

    build_exhaustive_index_gpu / query_exhaustive_index_gpu / query_exhaustive_index_gpu_self   (lib.rs:2813-2911)
    build_ivf_index_gpu / query_ivf_index_gpu / query_ivf_index_gpu_self                       (lib.rs:2913-3002)
    build_exhaustive_bf16_index ... / build_exhaustive_sq8_index ...                            (lib.rs:1702-1871)
    build_ivf_bf16_index ... / build_ivf_sq8_index ...                                          (lib.rs:2100-2290)

`faer::MatRef` becomes a 2-D numpy array (rows = samples), `(Vec<Vec<usize>>, Option<Vec<Vec<T>>>)`
becomes `(ids[nq, k] int64, dist[nq, k] float32 | None)` with `-1 / +inf` padding when fewer than k
neighbours exist, and `AnnSearchErrors` becomes `AnnSearchError` carrying the same variant name.

All arithmetic happens in the CUDA library; importing this module without the built shared
object, or calling it without a GPU, fails loudly -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ANNB200_LIB") or os.path.normpath(os.path.join(_HERE, "..", "..", "lib", "libannb200.so"))   # override: debugging builds

F32, BF16, SQ8 = 0, 1, 2
L2, COSINE, MANHATTAN = 0, 1, 2
PATH_AUTO, PATH_SIMT, PATH_TENSOR = 0, 1, 2

_STATUS_NAMES = {
    -1: "DimensionMismatch",
    -2: "DistanceNotSupported",
    -3: "TooFewSamplesForCentroids",
    -4: "InvalidArgument",
    -5: "Cuda",
    -6: "Nccl",
    -7: "OutOfMemory",
    -8: "Unsupported",
}


class AnnSearchError(RuntimeError):
    """Mirror of AnnSearchErrors (src/errors.rs): `.variant` names the enum variant."""

    def __init__(self, code: int, msg: str):
        self.code = code
        self.variant = _STATUS_NAMES.get(code, "Unknown")
        super().__init__(f"{self.variant}: {msg}")


class _Info(C.Structure):
    _fields_ = [("n", C.c_uint64), ("n_total", C.c_uint64), ("dim", C.c_uint32), ("nlist", C.c_uint32),
                ("dtype", C.c_int32), ("metric", C.c_int32), ("device", C.c_int32), ("is_ivf", C.c_int32),
                ("device_bytes", C.c_uint64), ("host_bytes", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C ann-search-rs_b200` "
                              "(or __graft_entry__.build()); annb200 has no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.annb_last_error.restype = C.c_char_p
        _lib.annb_parse_metric.argtypes = [C.c_char_p]
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        _lib.annb_flat_create.argtypes = [C.POINTER(vp), vp, u64, u32, i32, i32, vp, u64, i32]
        _lib.annb_flat_search.argtypes = [vp, vp, u64, u32, u32, vp, vp, vp]
        _lib.annb_flat_search_self.argtypes = [vp, u64, u64, u32, vp, vp, vp]
        _lib.annb_flat_search_dev.argtypes = [vp, vp, u64, u32, u32, vp, vp, vp, vp]
        _lib.annb_ivf_assign.argtypes = [vp, u64, u32, vp, vp, u32, i32, vp, i32]
        _lib.annb_kmeans_lloyd.argtypes = [vp, u64, u32, vp, u32, i32, u32, vp, i32]
        _lib.annb_kmeans_lloyd_balanced.argtypes = [vp, u64, u32, vp, u32, i32, u32, i32, u64, vp, vp, i32]
        _lib.annb_ivf_route_dev.argtypes = [vp, vp, u64, u32, u32, u32, vp, vp, u32, vp]
        _lib.annb_ivf_search_probes_dev.argtypes = [vp, vp, u64, u32, u32, u32, vp, vp, u32, vp, vp, vp, vp]
        _lib.annb_ivf_create.argtypes = [C.POINTER(vp), vp, vp, vp, vp, vp, vp, u64, u32, u32, i32, i32, vp, u32, u32, i32]
        _lib.annb_ivf_search.argtypes = [vp, vp, u64, u32, u32, u32, vp, vp, vp]
        _lib.annb_ivf_search_self.argtypes = [vp, u64, u64, u32, u32, i32, vp, vp, vp]
        _lib.annb_ivf_search_dev.argtypes = [vp, vp, u64, u32, u32, u32, vp, vp, vp, vp]
        _lib.annb_merge_topk_dev.argtypes = [vp, vp, u32, u64, u32, vp, vp, vp, vp]
        _lib.annb_flat_knn_graph.argtypes = [vp, u64, u64, u32, vp, vp, vp]
        _lib.annb_flat_search_shard_dev.argtypes = [vp, vp, u64, u32, u32, vp, vp, vp, vp]
        _lib.annb_ivf_search_probes_shard_dev.argtypes = [vp, vp, u64, u32, u32, u32, vp, vp, u32, vp, vp, vp, vp]
        _lib.annb_shard_check_dev.argtypes = [vp, vp, vp, u64, u32, C.POINTER(u32), vp]
        _lib.annb_shard_check_gathered_dev.argtypes = [vp, vp, u64, u64, u32, u32, vp, u64, u32, C.POINTER(u32), C.POINTER(u32), vp]
        _lib.annb_shard_check_gathered_async_dev.argtypes = [vp, vp, u64, u64, u32, u32, vp, u64, u32, vp, vp]
        _lib.annb_merge_check_shards_async_dev.argtypes = [vp, vp, u64, u64, u64, u32, u32, u64, u32, vp, vp, vp, vp]
        _lib.annb_ivf_validate.argtypes = [vp, vp, u64, u32, u32, C.POINTER(C.c_double)]
        _lib.annb_shard_refine_dev.argtypes = [vp, vp, u64, u32, u32, u32, vp, vp, u32, vp, vp, vp]
        _lib.annb_merge_shards_dev.argtypes = [vp, u64, u64, u32, u64, u32, vp, vp, vp, vp]
        _lib.annb_flat_create_multi.argtypes = [C.POINTER(vp), vp, u64, u32, i32, i32, vp, i32]
        _lib.annb_ivf_create_multi.argtypes = [C.POINTER(vp), vp, vp, vp, vp, vp, vp, u64, u32, u32, i32, i32, vp, vp, i32]
        _lib.annb_index_shard_count.argtypes = [vp, C.POINTER(u32)]
        _lib.annb_index_get_info.argtypes = [vp, C.POINTER(_Info)]
        _lib.annb_index_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
        _lib.annb_index_get_stat.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64)]
        _lib.annb_destroy.argtypes = [vp]
        _lib.annb_destroy.restype = None
        _lib.annb_device_count.argtypes = [C.POINTER(C.c_int)]
        _lib.annb_debug_fetch_tile.argtypes = [vp, vp]
    return _lib


def _check(code: int):
    if code != 0:
        raise AnnSearchError(code, lib().annb_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def device_count() -> int:
    c = C.c_int(0)
    _check(lib().annb_device_count(C.byref(c)))
    return c.value


def parse_ann_dist(s: str) -> Optional[int]:
    """src/utils/dist.rs:63-70."""
    m = lib().annb_parse_metric(s.encode())
    return None if m < 0 else m


def _metric_or_default(dist_metric: str) -> int:
    # src/lib.rs:274-277: unknown string -> warning + squared Euclidean
    m = parse_ann_dist(dist_metric)
    if m is None:
        print(f"  Unknown distance metric '{dist_metric}', defaulting to Euclidean")
        return L2
    return m


def matrix_to_flat(mat, device: int = 0) -> np.ndarray:
    """matrix_to_flat (src/utils/mod.rs:44-68) through annb_matrix_to_flat: any positively strided 2-D f32 view (e.g. a
    column-major / Fortran-order matrix, as faer stores them) -> contiguous row-major copy, gathered on the device."""
    a = np.asarray(mat)
    if a.ndim != 2 or a.dtype != np.float32:
        raise AnnSearchError(-4, "expected a 2-D float32 matrix (samples x features)")
    rs, cs = (st // 4 for st in a.strides)
    out = np.empty(a.shape, dtype=np.float32)
    f = lib().annb_matrix_to_flat
    f.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, C.c_int64, C.c_void_p, C.c_int32]
    f.restype = C.c_int32
    _check(f(C.c_void_p(a.ctypes.data), a.shape[0], a.shape[1], rs, cs, _ptr(out), device))
    return out


def _as_rowmajor_f32(mat) -> np.ndarray:
    """matrix_to_flat (src/utils/mod.rs:44-68): any-stride matrix -> contiguous row-major f32."""
    a = np.asarray(mat)
    if a.ndim != 2:
        raise AnnSearchError(-4, "expected a 2-D matrix (samples x features)")
    return np.ascontiguousarray(a, dtype=np.float32)


# --------------------------------------------------------------------------
# handles
# --------------------------------------------------------------------------
class _IndexBase:
    def __init__(self, handle: C.c_void_p):
        self._h = handle

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def info(self) -> _Info:
        i = _Info()
        _check(lib().annb_index_get_info(self._h, C.byref(i)))
        return i

    @property
    def n(self) -> int:
        return int(self.info().n)

    @property
    def dim(self) -> int:
        return int(self.info().dim)

    @property
    def shard_count(self) -> int:
        """Per-device shards behind the handle (1 unless it was built over a device list)."""
        c = C.c_uint32(0)
        _check(lib().annb_index_shard_count(self._h, C.byref(c)))
        return int(c.value)

    def memory_usage_bytes(self) -> Tuple[int, int]:
        """(ram, vram) as IvfIndexGpu::memory_usage_bytes (src/gpu/ivf_gpu.rs:590-604)."""
        i = self.info()
        return int(i.host_bytes), int(i.device_bytes)

    def set_option(self, key: str, value: int):
        _check(lib().annb_index_set_option(self._h, key.encode(), C.c_int64(value)))

    def get_stat(self, key: str) -> int:
        v = C.c_int64(0)
        _check(lib().annb_index_get_stat(self._h, key.encode(), C.byref(v)))
        return int(v.value)

    def debug_fetch_tile(self) -> np.ndarray:
        out = np.empty((128, 128), dtype=np.float32)
        _check(lib().annb_debug_fetch_tile(self._h, _ptr(out)))
        return out

    def uncertified_queries(self) -> np.ndarray:
        """Batch-relative numbers of the queries of the last tensor-path batch that failed the coverage certificate."""
        f = lib().annb_debug_fetch_uncertified
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        cnt = C.c_uint32(0)
        buf = np.zeros(1 << 16, dtype=np.uint32)
        _check(f(self._h, _ptr(buf), buf.size, C.byref(cnt)))
        return buf[:min(int(cnt.value), buf.size)].copy()

    def cert_eps(self) -> float:
        """Error bound the last tensor-path certificate assumed."""
        return float(np.array([self.get_stat("cert_eps_bits")], dtype=np.uint32).view(np.float32)[0])

    def close(self):
        if self._h is not None:
            lib().annb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ExhaustiveIndexB200(_IndexBase):
    """Resident flat index: ExhaustiveIndexGpu (src/gpu/exhaustive_gpu.rs:17-33) and its BF16 / SQ8 twins."""

    @classmethod
    def new(cls, data, metric: int, dtype: int = F32, device=0, id_base: int = 0, sq8_scales=None):
        """`device`: one ordinal, or a list of ordinals -- the rows are then sharded over those GPUs behind one handle
        (annb_flat_create_multi) and every query method below works unchanged."""
        x = _as_rowmajor_f32(data)
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*[int(d) for d in device])
            _check(lib().annb_flat_create_multi(C.byref(h), _ptr(x), x.shape[0], x.shape[1], dtype, metric, devs, len(device)))
            return cls(h)
        sc = None if sq8_scales is None else np.ascontiguousarray(sq8_scales, dtype=np.float32)
        _check(lib().annb_flat_create(C.byref(h), _ptr(x), x.shape[0], x.shape[1], dtype, metric, _ptr(sc), id_base, device))
        return cls(h)

    def query_batch(self, query_mat, k: int, return_dist: bool = True):
        q = _as_rowmajor_f32(query_mat)
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.uint64)
        dist = np.empty((nq, k), dtype=np.float32) if return_dist else None
        cnt = np.empty(nq, dtype=np.uint32)
        _check(lib().annb_flat_search(self._h, _ptr(q), nq, q.shape[1], k, _ptr(ids), _ptr(dist), _ptr(cnt)))
        return ids.view(np.int64), dist, cnt

    def generate_knn(self, k: int, return_dist: bool = True, row_begin: int = 0, row_end: Optional[int] = None):
        row_end = self.n if row_end is None else row_end
        nq = row_end - row_begin
        ids = np.empty((nq, k), dtype=np.uint64)
        dist = np.empty((nq, k), dtype=np.float32) if return_dist else None
        cnt = np.empty(nq, dtype=np.uint32)
        _check(lib().annb_flat_search_self(self._h, row_begin, row_end, k, _ptr(ids), _ptr(dist), _ptr(cnt)))
        return ids.view(np.int64), dist, cnt


    def knn_graph(self, k: int, row_begin: int = 0, row_end: Optional[int] = None):
        """annb_flat_knn_graph: (pid [rows, k] int64, dist [rows, k] float32, counts [rows]) -- every row's k nearest OTHER
        rows, ascending, padded with (SENTINEL_PID, f32::MAX)."""
        row_end = self.n if row_end is None else row_end
        nq = row_end - row_begin
        pid = np.empty((nq, k), dtype=np.uint64)
        dist = np.empty((nq, k), dtype=np.float32)
        cnt = np.empty(nq, dtype=np.uint32)
        _check(lib().annb_flat_knn_graph(self._h, row_begin, row_end, k, _ptr(pid), _ptr(dist), _ptr(cnt)))
        return pid.view(np.int64), dist, cnt


class IvfIndexB200(_IndexBase):
    """Resident IVF index: IvfIndexGpu (src/gpu/ivf_gpu.rs:153-181) and its BF16 / SQ8 twins."""

    def __init__(self, handle, parts: Optional[dict] = None):
        super().__init__(handle)
        self.parts = parts or {}

    @classmethod
    def from_parts(cls, vectors, centroids, offsets, original_ids, dtype: int, metric: int, norms=None,
                   centroid_norms=None, sq8_scales=None, list_begin: int = 0, list_end: Optional[int] = None,
                   device=0, n_total: Optional[int] = None):
        """annb_ivf_create: the contents of the reference's IvfIndex struct, already in list order.  `device` may be a
        list of ordinals: the inverted lists are then sharded over those GPUs behind one handle (annb_ivf_create_multi)."""
        want = {F32: np.float32, BF16: np.uint16, SQ8: np.int8}[dtype]
        v = np.ascontiguousarray(vectors, dtype=want)
        cent = np.ascontiguousarray(centroids, dtype=np.float32)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        oid = np.ascontiguousarray(original_ids, dtype=np.uint64)
        nlist = cent.shape[0]
        list_end = nlist if list_end is None else list_end
        n = int(off[-1]) if n_total is None else n_total
        nr = None
        if norms is not None:
            nr = np.ascontiguousarray(norms, dtype=np.int32 if dtype == SQ8 else np.float32)
        cn = None if centroid_norms is None else np.ascontiguousarray(centroid_norms, dtype=np.float32)
        sc = None if sq8_scales is None else np.ascontiguousarray(sq8_scales, dtype=np.float32)
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*[int(d) for d in device])
            _check(lib().annb_ivf_create_multi(C.byref(h), _ptr(v), _ptr(nr), _ptr(cent), _ptr(cn), _ptr(off), _ptr(oid), n,
                                               cent.shape[1], nlist, dtype, metric, _ptr(sc), devs, len(device)))
            return cls(h)
        _check(lib().annb_ivf_create(C.byref(h), _ptr(v), _ptr(nr), _ptr(cent), _ptr(cn), _ptr(off), _ptr(oid), n,
                                     cent.shape[1], nlist, dtype, metric, _ptr(sc), list_begin, list_end, device))
        return cls(h)

    def query_batch(self, query_mat, k: int, nprobe: Optional[int] = None, return_dist: bool = True):
        q = _as_rowmajor_f32(query_mat)
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.uint64)
        dist = np.empty((nq, k), dtype=np.float32) if return_dist else None
        cnt = np.empty(nq, dtype=np.uint32)
        _check(lib().annb_ivf_search(self._h, _ptr(q), nq, q.shape[1], k, nprobe or 0, _ptr(ids), _ptr(dist), _ptr(cnt)))
        return ids.view(np.int64), dist, cnt

    def generate_knn(self, k: int, nprobe: Optional[int] = None, return_dist: bool = True, pos_begin: int = 0,
                     pos_end: Optional[int] = None, scatter: Optional[bool] = None):
        n = self.n
        pos_end = n if pos_end is None else pos_end
        full = pos_begin == 0 and pos_end == n
        scatter = full if scatter is None else scatter
        rows = n if scatter else pos_end - pos_begin
        ids = np.full((rows, k), np.iinfo(np.uint64).max, dtype=np.uint64)
        dist = np.full((rows, k), np.inf, dtype=np.float32) if return_dist else None
        cnt = np.zeros(rows, dtype=np.uint32)
        _check(lib().annb_ivf_search_self(self._h, pos_begin, pos_end, k, nprobe or 0, 1 if scatter else 0, _ptr(ids), _ptr(dist),
                                          _ptr(cnt)))
        return ids.view(np.int64), dist, cnt


    def validate_index(self, k: int, seed: int = 42, no_samples: Optional[int] = None, positions=None) -> float:
        """KnnValidation::validate_index (src/utils/mod.rs:210-242, src/cpu/ivf.rs:496-523): recall@k of the index against an
        exhaustive search over its own vectors on `no_samples` (default min(1000, n)) stored vectors drawn with replacement.
        numpy's PCG64 stands in for StdRng (or pass the internal positions yourself); unsharded f32 indices only."""
        n = self.n
        ns = min(1000 if no_samples is None else no_samples, n)
        if positions is None:
            positions = np.random.Generator(np.random.PCG64(seed)).integers(0, n, size=ns, dtype=np.uint64)
        pos = np.ascontiguousarray(positions, dtype=np.uint64)
        out = C.c_double(0.0)
        _check(lib().annb_ivf_validate(self._h, _ptr(pos), pos.size, k, 0, C.byref(out)))
        return float(out.value)


# --------------------------------------------------------------------------
# host-side build steps of the IVF constructors (integer / byte work the reference also does on the CPU)
# --------------------------------------------------------------------------
def encode_bf16(x: np.ndarray) -> np.ndarray:
    """encode_bf16_quantisation (src/quantised/quantisers.rs:31-38): round-to-nearest-even."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    nan = (b & 0x7FFFFFFF) > 0x7F800000
    up = ((b & 0x8000) != 0) & ((b & 0x17FFF) != 0)
    out = (b >> 16) + up.astype(np.uint32)
    out = np.where(nan, (b >> 16) | 0x40, out)
    return out.astype(np.uint16)


def sq8_train(x: np.ndarray) -> np.ndarray:
    """ScalarQuantiser::train (src/quantised/quantisers.rs:123-146)."""
    mx = np.abs(np.asarray(x, dtype=np.float32)).max(axis=0)
    return np.where(mx <= 0, np.float32(1.0), mx / np.float32(128.0)).astype(np.float32)


def sq8_encode(x: np.ndarray, scales: np.ndarray) -> np.ndarray:
    """ScalarQuantiser::encode (src/quantised/quantisers.rs:148-165)."""
    scaled = np.asarray(x, dtype=np.float32) / np.asarray(scales, dtype=np.float32)[None, :]
    sg = np.where(np.signbit(scaled), np.float32(-1.0), np.float32(1.0))
    r = (scaled + np.float32(0.5) * sg).astype(np.float32)
    r = np.clip(r, np.float32(-128.0), np.float32(127.0))
    return np.trunc(np.nan_to_num(r, nan=0.0)).astype(np.int8)


def build_csr_layout(assignments: np.ndarray, nlist: int):
    """build_csr_layout (src/utils/k_means_utils.rs:2955-2980): stable counting sort."""
    a = np.asarray(assignments, dtype=np.int64)
    counts = np.bincount(a, minlength=nlist)
    offsets = np.zeros(nlist + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(counts)
    return np.argsort(a, kind="stable").astype(np.uint64), offsets


def ivf_assign(data, centroids, metric: int, centroid_norms=None, device: int = 0) -> np.ndarray:
    x = _as_rowmajor_f32(data)
    c = _as_rowmajor_f32(centroids)
    cn = None if centroid_norms is None else np.ascontiguousarray(centroid_norms, dtype=np.float32)
    out = np.empty(x.shape[0], dtype=np.uint32)
    _check(lib().annb_ivf_assign(_ptr(x), x.shape[0], x.shape[1], _ptr(c), _ptr(cn), c.shape[0], metric, _ptr(out), device))
    return out


# --------------------------------------------------------------------------
# free functions mirroring src/lib.rs
# --------------------------------------------------------------------------
def _finish(res, return_dist):
    ids, dist, _ = res
    return ids, (dist if return_dist else None)


def build_exhaustive_index_gpu(mat, dist_metric: str = "euclidean", device=0) -> ExhaustiveIndexB200:
    """src/lib.rs:2813-2840.  `device` (the reference's R::Device argument): a GPU ordinal, or a list of ordinals to
    shard the rows over several GPUs of the box."""
    return ExhaustiveIndexB200.new(mat, _metric_or_default(dist_metric), F32, device)


def query_exhaustive_index_gpu(query_mat, index: ExhaustiveIndexB200, k: int, return_dist: bool = True, verbose: bool = False):
    """src/lib.rs:2842-2875."""
    return _finish(index.query_batch(query_mat, k, return_dist), return_dist)


def query_exhaustive_index_gpu_self(index: ExhaustiveIndexB200, k: int, return_dist: bool = True, verbose: bool = False):
    """src/lib.rs:2877-2911."""
    return _finish(index.generate_knn(k, return_dist), return_dist)


def build_exhaustive_bf16_index(mat, dist_metric: str = "euclidean", device: int = 0) -> ExhaustiveIndexB200:
    """src/lib.rs:1702-1731."""
    return ExhaustiveIndexB200.new(mat, _metric_or_default(dist_metric), BF16, device)


def build_exhaustive_sq8_index(mat, dist_metric: str = "euclidean", device: int = 0) -> ExhaustiveIndexB200:
    """src/lib.rs:1787-1816."""
    return ExhaustiveIndexB200.new(mat, _metric_or_default(dist_metric), SQ8, device)


query_exhaustive_bf16_index = query_exhaustive_index_gpu       # src/lib.rs:1733-1760
query_exhaustive_bf16_self = query_exhaustive_index_gpu_self   # src/lib.rs:1762-1785
query_exhaustive_sq8_index = query_exhaustive_index_gpu        # src/lib.rs:1818-1845
query_exhaustive_sq8_self = query_exhaustive_index_gpu_self    # src/lib.rs:1847-1871


SENTINEL_PID = (2 ** 32 - 1) >> 1          # src/utils/nndescent_utils.rs:25


class KnnGraphGpu:
    """Mirror of KnnGraphGpu<T> (src/gpu/nndescent_gpu.rs:2418-2446), the hand-off struct between the GPU kNN-graph
    builders and their consumers (build_nsg_from_gpu_knn, src/lib.rs:3330-3345; raw kNN extraction).  Same fields; the
    flat `Vec<(usize, T)>` graph is kept as two [n, k] arrays (`pid`, `dist`), `knn_graph` yields the reference's pairs."""

    def __init__(self, vectors_flat, dim, n, k, norms, metric, pid, dist, converged=True):
        self.vectors_flat, self.dim, self.n, self.k = vectors_flat, dim, n, k
        self.norms, self.metric, self.pid, self.dist, self.converged = norms, metric, pid, dist, converged

    @property
    def knn_graph(self):
        return list(zip(self.pid.reshape(-1).tolist(), self.dist.reshape(-1).tolist()))

    def check_contract(self):
        """What NsgIndex::build_from_knn (src/cpu/nsg.rs:744-775) relies on: n * k entries, rows ascending by distance,
        no self edge, ids < n or the sentinel, sentinels (and only sentinels) at the tail with T::MAX distances."""
        pid, dist = self.pid, self.dist
        assert pid.shape == (self.n, self.k) and dist.shape == (self.n, self.k)
        real = pid != SENTINEL_PID
        assert ((pid < self.n) | ~real).all() and (pid >= 0).all()
        assert (pid != np.arange(self.n)[:, None]).all(), "self edge"
        assert (real[:, :-1] | ~real[:, 1:]).all(), "sentinels must be trailing"
        assert (dist[~real] == np.finfo(np.float32).max).all()
        assert (np.diff(dist, axis=1) >= 0).all(), "rows must ascend by distance"
        return True


def build_knn_graph_gpu(mat, dist_metric: str = "euclidean", k: Optional[int] = None, build_k=None, max_iters=None, n_trees=None, delta=None,
                        rho=None, refine_knn=None, seed: int = 42, verbose: bool = False, device=0) -> KnnGraphGpu:
    """src/lib.rs:3201-3228 with the graph computed exactly: the exhaustive self search of the flat index (BASELINE
    configs[4]) instead of NN-Descent, so the NN-Descent tuning arguments are accepted and ignored and `converged` is
    always true.  `k` defaults to 30 as in the reference.  `device`: an ordinal or a list of ordinals."""
    metric = _metric_or_default(dist_metric)
    if metric == MANHATTAN:
        raise AnnSearchError(-2, "Manhattan distance is not supported by the GPU kNN-graph builder")
    x = _as_rowmajor_f32(mat)
    n, dim = x.shape
    k = 30 if k is None else int(k)
    ix = ExhaustiveIndexB200.new(x, metric, F32, device)
    try:
        pid, dist, _ = ix.knn_graph(k)
    finally:
        ix.close()
    norms = ref_row_norms(x) if metric == COSINE else np.zeros(0, dtype=np.float32)
    return KnnGraphGpu(x.reshape(-1), dim, n, k, norms, metric, pid, dist, True)


def ref_row_norms(x: np.ndarray) -> np.ndarray:
    """calculate_l2_norm per row in the AVX2 lane order (src/utils/dist.rs:2339-2360): 8 lane accumulators over
    8-element chunks (separate multiply and add), wide's reduce_add tree, sequential tail, sqrt."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, dim = x.shape
    chunks = dim // 8
    acc = np.zeros((n, 8), dtype=np.float32)
    for c in range(chunks):
        v = x[:, c * 8:(c + 1) * 8]
        acc += v * v
    s = acc[:, :4] + acc[:, 4:]
    tot = (s[:, 0] + s[:, 2]) + (s[:, 1] + s[:, 3])
    for e in range(chunks * 8, dim):
        tot = tot + x[:, e] * x[:, e]
    return np.sqrt(tot).astype(np.float32)


def seq_row_norms(x: np.ndarray) -> np.ndarray:
    """Sequential-fold norms (centroid norms, src/cpu/ivf.rs:193-206)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    s = np.zeros(x.shape[0], dtype=np.float32)
    for e in range(x.shape[1]):
        s = s + x[:, e] * x[:, e]
    return np.sqrt(s).astype(np.float32)


def normalise_rows(x: np.ndarray) -> np.ndarray:
    """normalise_vector per row (src/utils/dist.rs:5336-5344)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    nrm = ref_row_norms(x)
    safe = np.where(nrm > 0, nrm, np.float32(1.0)).astype(np.float32)
    return np.where((nrm > 0)[:, None], x / safe[:, None], x).astype(np.float32)


def kmeans_lloyd(train: np.ndarray, init_centroids: np.ndarray, metric: int, max_iters: int = 30, device: int = 0, balanced: bool = False,
                 seed: int = 42, return_moves: bool = False):
    """Lloyd iterations of train_centroids on the device (annb_kmeans_lloyd[_balanced]): parallel_lloyd
    (src/utils/k_means_utils.rs:1572-1700) from the given initial centroids, optionally with the balancing hook
    (adjust_centers, :979-1030).  Returns (centroids, updates done[, centroid moves])."""
    train = _as_rowmajor_f32(train)
    cent = _as_rowmajor_f32(init_centroids).copy()
    if train.shape[1] != cent.shape[1]:
        raise AnnSearchError(-1, f"training data has dim {train.shape[1]}, centroids {cent.shape[1]}")
    it = C.c_uint32(0)
    moves = C.c_uint64(0)
    _check(lib().annb_kmeans_lloyd_balanced(_ptr(train), train.shape[0], train.shape[1], _ptr(cent), cent.shape[0], metric, max_iters,
                                            1 if balanced else 0, seed, C.byref(it), C.byref(moves), device))
    return (cent, int(it.value), int(moves.value)) if return_moves else (cent, int(it.value))


def _ref_euclid_rows(x: np.ndarray, c: np.ndarray) -> np.ndarray:
    """euclidean_distance_static of every row of x to the vector c in the AVX2 lane order (src/utils/dist.rs:306-330):
    8 lane accumulators over 8-element chunks (separate multiply and add), wide's reduce_add tree, sequential tail."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    c = np.ascontiguousarray(c, dtype=np.float32)
    n, dim = x.shape
    chunks = dim // 8
    acc = np.zeros((n, 8), dtype=np.float32)
    for ch in range(chunks):
        d = x[:, ch * 8:(ch + 1) * 8] - c[None, ch * 8:(ch + 1) * 8]
        acc += d * d
    s = acc[:, :4] + acc[:, 4:]
    tot = (s[:, 0] + s[:, 2]) + (s[:, 1] + s[:, 3])
    for e in range(chunks * 8, dim):
        d = x[:, e] - c[e]
        tot = tot + d * d
    return tot.astype(np.float32)


def _pick_by_cumsum(dist: np.ndarray, u: float) -> int:
    """`threshold = u * total; first idx with cumsum >= threshold` with the reference's sequential f64 sums
    (src/utils/k_means_utils.rs:580-592, 495-506); the last index if rounding leaves the threshold above the sum."""
    cs = np.cumsum(dist.astype(np.float64))
    return int(min(np.searchsorted(cs, u * cs[-1], side="left"), dist.size - 1))


def fast_random_init(train: np.ndarray, k: int, rng) -> np.ndarray:
    """fast_random_init (src/utils/k_means_utils.rs:612-626): k distinct random rows.  `rng`: numpy Generator standing in
    for rand's StdRng shuffle (the stream itself cannot be reproduced without the crate)."""
    return np.ascontiguousarray(train[rng.permutation(train.shape[0])[:k]])


def kmeans_parallel_init(train: np.ndarray, k: int, metric: int, rng, device: int = 0) -> np.ndarray:
    """k-means|| seeding (kmeans_parallel_init + weighted_kmeans_plus_plus, src/utils/k_means_utils.rs:435-596) for the
    squared-Euclidean metric: ln(k) + 1 rounds, each a D^2 pass of every training row against the candidates so far --
    on the GPU, as a k = 1 search of a flat index holding the candidates (the reference's min_distance_to_centroids
    arithmetic, bit for bit) -- followed by 2k draws proportional to D^2 (sequential f64 cumulative sums on the host, as the
    reference); the oversampled candidates are then reduced to k with the reference's k-means++ walk.  The uniform
    draws come from `rng` in the reference's order (first index, then per round the 2k thresholds, then the k-means++
    draws), so a caller that can replay StdRng's stream reproduces the reference's centroids."""
    x = _as_rowmajor_f32(train)
    n, dim = x.shape
    if metric != L2:
        raise AnnSearchError(-8, "k-means|| seeding on the device is restated for the squared-Euclidean metric")
    rounds = int(math.log(k) + 1.0)
    cand_rows = [int(rng.integers(0, n))]
    for _ in range(rounds):
        ix = ExhaustiveIndexB200.new(x[cand_rows], L2, F32, device)
        try:
            _, d, _ = ix.query_batch(x, 1)
        finally:
            ix.close()
        d = d[:, 0]
        for _ in range(2 * k):
            cand_rows.append(_pick_by_cumsum(d, float(rng.random())))
    cand = np.ascontiguousarray(x[cand_rows])
    m = cand.shape[0]
    if m <= k:
        return cand
    chosen = [int(rng.integers(0, m))]
    dist = np.full(m, np.inf, dtype=np.float32)
    for _ in range(1, k):
        dist = np.minimum(dist, _ref_euclid_rows(cand, cand[chosen[-1]]))
        chosen.append(_pick_by_cumsum(dist, float(rng.random())))
    return np.ascontiguousarray(cand[chosen])


def train_centroids(train: np.ndarray, nlist: int, metric: int, iters: int = 30, device: int = 0, seed: int = 42, init: Optional[str] = None,
                    balanced: bool = False) -> np.ndarray:
    """train_centroids (src/utils/k_means_utils.rs:2771-2938) on the device: seeding as resolve_init chooses it (:233-241:
    more than 200 centroids -> k distinct random rows, otherwise k-means||; cosine always seeds with random rows here), then
    the Lloyd loop of annb_kmeans_lloyd[_balanced].  numpy's PCG64 stands in for rand's StdRng: the algorithm is the
    reference's, the individual draws are not.  The Hamerly / GEMM variants of the loop (:261-284) compute the same
    assignments with bounds / a third-party GEMM and are not restated: every path runs the direct-assignment loop."""
    train = _as_rowmajor_f32(train)
    n, dim = train.shape
    if n < nlist:
        raise AnnSearchError(-3, f"{n} training samples for {nlist} centroids")
    rng = np.random.Generator(np.random.PCG64(seed))
    how = init or ("random" if (nlist > 200 or metric != L2) else "kmeans||")
    if how == "random":
        cent0 = fast_random_init(train, nlist, rng)
    elif how == "kmeans||":
        cent0 = kmeans_parallel_init(train, nlist, metric, rng, device)
    elif how == "spaced":
        cent0 = train[(np.arange(nlist, dtype=np.int64) * n) // nlist].copy()
    else:
        raise AnnSearchError(-4, f"unknown k-means init '{how}'")
    return kmeans_lloyd(train, cent0, metric, max(1, iters), device, balanced=balanced, seed=seed)[0]


def train_centroids_lloyd(train: np.ndarray, nlist: int, metric: int, iters: int = 30, device: int = 0) -> np.ndarray:
    """Kept for callers of round 1: Lloyd from evenly spaced training rows."""
    return train_centroids(train, nlist, metric, iters, device, init="spaced")


def build_ivf_host_parts(mat, centroids, metric: int, dtype: int, train_rows=None, device: int = 0) -> dict:
    """The steps of IvfIndex::build after training (src/cpu/ivf.rs:164-249), IvfIndexBf16::build
    (src/quantised/ivf_bf16.rs:150-254) and IvfSq8Index::build (src/quantised/ivf_sq8.rs:158-284):
    norms, centroid norms, GPU assignment, CSR layout, list-order permutation, quantisation."""
    x = _as_rowmajor_f32(mat)
    cent = _as_rowmajor_f32(centroids).copy()
    n, dim = x.shape
    nlist = cent.shape[0]
    norms = cnorms = scales = norms_i = None
    if dtype == SQ8:
        if metric == COSINE:
            x = normalise_rows(x)
            cent = normalise_rows(cent)
        tr = x if train_rows is None else x[np.asarray(train_rows)]
        scales = sq8_train(tr)                                   # codebook from the training sample (ivf_sq8.rs:211)
        assign = ivf_assign(x, cent, metric, np.ones(nlist, dtype=np.float32), device)
    else:
        if metric == COSINE:
            norms = ref_row_norms(x)
            cnorms = seq_row_norms(cent)
        assign = ivf_assign(x, cent, metric, cnorms, device)
    new_to_old, offsets = build_csr_layout(assign, nlist)
    order = new_to_old.astype(np.int64)
    if dtype == F32:
        vec = np.ascontiguousarray(x[order])
    elif dtype == BF16:
        vec = np.ascontiguousarray(encode_bf16(x)[order])
    else:
        vec = np.ascontiguousarray(sq8_encode(x, scales)[order])
        if metric == COSINE:
            norms_i = (vec.astype(np.int32) ** 2).sum(axis=1).astype(np.int32)
    return dict(vectors=vec, centroids=cent, offsets=offsets, original_ids=new_to_old, dtype=dtype, metric=metric,
                norms=(norms[order] if norms is not None else norms_i), centroid_norms=cnorms, sq8_scales=scales)


def _first_device(device) -> int:
    return int(device[0]) if isinstance(device, (list, tuple)) else int(device)


def _build_ivf(mat, nlist, centroids, dist_metric, dtype, seed, device, verbose, kmeans_iters=30, kmeans_init=None, balanced=False) -> IvfIndexB200:
    metric = _metric_or_default(dist_metric)
    if metric == MANHATTAN:
        raise AnnSearchError(-2, "Manhattan distance is not supported by the IVF indices")
    x = _as_rowmajor_f32(mat)
    n, dim = x.shape
    if nlist is None:
        nlist = default_nlist(n) if centroids is None else int(np.asarray(centroids).shape[0])
    nlist = max(int(nlist), 1)
    n_train = max(min(256 * nlist, 250_000, n), 1)                 # src/cpu/ivf.rs:174
    rng = np.random.Generator(np.random.PCG64(seed))               # stand-in for StdRng (sample_vectors)
    train_rows = rng.permutation(n)[:n_train]
    if centroids is None:
        xt = x[train_rows]
        if dtype == SQ8 and metric == COSINE:
            xt = normalise_rows(xt)
        if verbose:
            print(f"  Generating IVF index with {nlist} Voronoi cells.")
        centroids = train_centroids(xt, nlist, metric, kmeans_iters, _first_device(device), seed=seed, init=kmeans_init, balanced=balanced)
    parts = build_ivf_host_parts(x, centroids, metric, dtype, train_rows, _first_device(device))
    ix = IvfIndexB200.from_parts(device=device, **parts)
    ix.parts = parts
    return ix


def build_ivf_index_gpu(mat, nlist: Optional[int] = None, k_means_params=None, dist_metric: str = "euclidean", seed: int = 42,
                        verbose: bool = False, device: int = 0, centroids=None) -> IvfIndexB200:
    """src/lib.rs:2913-2947.  `k_means_params` mirrors KMeansTrainingParams (src/utils/k_means_utils.rs:286-345) as a dict:
    {"iters": int, "init": "random" | "kmeans||" | None, "balanced": bool}; `centroids` short-circuits training."""
    kp = k_means_params or {}
    return _build_ivf(mat, nlist, centroids, dist_metric, F32, seed, device, verbose, kp.get("iters", 30), kp.get("init"), bool(kp.get("balanced", False)))


def build_ivf_bf16_index(mat, nlist: Optional[int] = None, k_means_params=None, dist_metric: str = "euclidean", seed: int = 42,
                         verbose: bool = False, device: int = 0, centroids=None) -> IvfIndexB200:
    """src/lib.rs:2100-2140."""
    kp = k_means_params or {}
    return _build_ivf(mat, nlist, centroids, dist_metric, BF16, seed, device, verbose, kp.get("iters", 30), kp.get("init"), bool(kp.get("balanced", False)))


def build_ivf_sq8_index(mat, nlist: Optional[int] = None, k_means_params=None, dist_metric: str = "euclidean", seed: int = 42,
                        verbose: bool = False, device: int = 0, centroids=None) -> IvfIndexB200:
    """src/lib.rs:2196-2236."""
    kp = k_means_params or {}
    return _build_ivf(mat, nlist, centroids, dist_metric, SQ8, seed, device, verbose, kp.get("iters", 30), kp.get("init"), bool(kp.get("balanced", False)))


def query_ivf_index_gpu(query_mat, index: IvfIndexB200, k: int, nprobe: Optional[int] = None, nquery: Optional[int] = None,
                        return_dist: bool = True, verbose: bool = False):
    """src/lib.rs:2949-2987.  `nquery` (the reference's batch size) is accepted and ignored: batching is internal."""
    return _finish(index.query_batch(query_mat, k, nprobe, return_dist), return_dist)


def query_ivf_index_gpu_self(index: IvfIndexB200, k: int, nprobe: Optional[int] = None, nquery: Optional[int] = None,
                             return_dist: bool = True, verbose: bool = False):
    """src/lib.rs:2989-3002."""
    return _finish(index.generate_knn(k, nprobe, return_dist), return_dist)


def default_nlist(n: int) -> int:
    """src/cpu/ivf.rs:172."""
    return max(1, int(np.float32(n) ** np.float32(0.5)))


def default_nprobe(nlist: int) -> int:
    """src/cpu/ivf.rs:345-347."""
    return max(1, int(math.sqrt(nlist)))


query_ivf_bf16_index = query_ivf_index_gpu          # src/lib.rs:2142-2170
query_ivf_bf16_self = query_ivf_index_gpu_self      # src/lib.rs:2172-2194
query_ivf_sq8_index = query_ivf_index_gpu           # src/lib.rs:2238-2266
query_ivf_sq8_self = query_ivf_index_gpu_self       # src/lib.rs:2268-2290
