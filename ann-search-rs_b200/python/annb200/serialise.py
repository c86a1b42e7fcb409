"""Reader / writer for the crate's saved indices (`IndexIo`, src/serialise/mod.rs:33-335), kinds "exhaustive", "ivf" and
the quantised twins "exhaustive_bf16", "exhaustive_sq8", "ivf_bf16", "ivf_sq8", so that an index built and saved by the CPU
crate can be served from a B200 and the other way round.

File layout (src/serialise/mod.rs:28-106, 240-330): a directory holding `index.bin` =
    8 magic bytes b"ANNSRS\\0\\0" | u32 LE format version (2) | u8 float width | u8 kind length | kind tag |
    bincode 2 payload of the serde-derived struct, `bincode::config::standard()`: little-endian, variable-length
    integers (one byte below 251, else marker 251 / 252 / 253 + u16 / u32 / u64), f32 as 4 raw bytes, `Vec<T>` as
    length + elements, unit enum variants as their u32 index, fields in declaration order:
      ExhaustiveIndex<T> (src/cpu/exhaustive.rs:18-32): vectors_flat, dim, n, norms, metric
      IvfIndex<T>        (src/cpu/ivf.rs:24-48): vectors_flat, dim, n, norms, metric, centroids, centroids_norm,
                         all_indices, offsets, nlist, original_ids
      ExhaustiveIndexBf16<T> (src/quantised/exhaustive_bf16.rs:24-39): vectors_flat (Vec<bf16>), dim, n, norms, metric, _phantom
      ExhaustiveSq8Index<T>  (src/quantised/exhaustive_sq8.rs:38-45): quantised_vectors (Vec<i8>), quantised_norms (Vec<i32>), dim, n,
                         metric, codebook
      IvfIndexBf16<T>    (src/quantised/ivf_bf16.rs:24-49): as IvfIndex with vectors_flat: Vec<bf16>
      IvfSq8Index<T>     (src/quantised/ivf_sq8.rs:27-52): quantised_vectors, quantised_norms, dim, n, metric, centroids, centroids_norm,
                         all_indices, offsets, codebook, nlist, original_ids
      ScalarQuantiser<T> (src/quantised/quantisers.rs:103-107): scales
      Dist               (src/utils/dist.rs:29-37): SquaredEuclidean = 0, Cosine = 1, Manhattan = 2
    Element encodings under the standard configuration: `half::bf16` is a serde newtype over u16 and therefore a
    variable-length integer like any u16 (one byte below 251, else 0xFB + 2 bytes); i8 is one raw byte; i32 is zig-zag
    mapped ((v << 1) ^ (v >> 31)) and then variable-length; PhantomData writes nothing.

PINNING STATUS.  The header and every error path below restate the reference's own tests (src/serialise/mod.rs:1514-1640).
The payload encoding is restated from bincode 2's published specification: bincode is a third-party crate (2.x, Cargo.lock)
that is not under /root/reference, there is no saved index in the reference tree and no Rust toolchain in this image, so
the payload bytes are **unpinned** -- checked here only by round trips and against the specification's own boundary cases.
"""
import ctypes as C
import os

import numpy as np

from . import lib

MAGIC = b"ANNSRS\0\0"
FORMAT_VERSION = 2
INDEX_FILE = "index.bin"
DIST_NAMES = ("SquaredEuclidean", "Cosine", "Manhattan")


class SerialiseError(RuntimeError):
    """Mirror of the serialisation variants of AnnSearchErrors (src/errors.rs): `.variant` names the enum variant."""

    def __init__(self, variant: str, msg: str, **fields):
        self.variant = variant
        self.fields = fields
        super().__init__(f"{variant}: {msg}")


# ----------------------------------------------------------------------------------------------- primitives
def _varint(v: int) -> bytes:
    if v < 0:
        raise SerialiseError("EncodeError", f"negative length / index {v}")
    if v < 251:
        return bytes([v])
    if v <= 0xFFFF:
        return b"\xfb" + v.to_bytes(2, "little")
    if v <= 0xFFFFFFFF:
        return b"\xfc" + v.to_bytes(4, "little")
    return b"\xfd" + v.to_bytes(8, "little")


def encode_usize_vec(values) -> bytes:
    """Vec<usize>: length, then one varint per element (encoded by the C helper: millions of ids)."""
    a = np.ascontiguousarray(values, dtype=np.uint64)
    out = np.empty(a.size * 9 + 16, dtype=np.uint8)
    f = lib().annb_varint_encode_u64
    f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    f.restype = C.c_int64
    n = f(C.c_void_p(a.ctypes.data), a.size, C.c_void_p(out.ctypes.data), out.size)
    if n < 0:
        raise SerialiseError("EncodeError", "varint encoding failed")
    return _varint(a.size) + out[:n].tobytes()


class _Reader:
    def __init__(self, buf: bytes, pos: int):
        self.buf, self.pos = buf, pos

    def _need(self, n: int):
        if self.pos + n > len(self.buf):
            raise SerialiseError("DecodeError", "unexpected end of the payload")

    def varint(self) -> int:
        self._need(1)
        m = self.buf[self.pos]
        self.pos += 1
        if m < 251:
            return m
        if m > 253:
            raise SerialiseError("DecodeError", f"integer marker {m} in a 64-bit field")
        n = {251: 2, 252: 4, 253: 8}[m]
        self._need(n)
        v = int.from_bytes(self.buf[self.pos:self.pos + n], "little")
        self.pos += n
        return v

    def f32_vec(self) -> np.ndarray:
        n = self.varint()
        self._need(4 * n)
        a = np.frombuffer(self.buf, dtype="<f4", count=n, offset=self.pos).copy()
        self.pos += 4 * n
        return a

    def usize_vec(self) -> np.ndarray:
        n = self.varint()
        if n > len(self.buf) - self.pos:      # every element takes at least one byte: reject absurd counts before allocating
            raise SerialiseError("DecodeError", "unexpected end of the payload inside an index list")
        out = np.empty(n, dtype=np.uint64)
        f = lib().annb_varint_decode_u64
        f.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p]
        f.restype = C.c_int64
        view = self.buf[self.pos:]
        used = f(view, len(view), n, C.c_void_p(out.ctypes.data))
        if used < 0:
            raise SerialiseError("DecodeError", "unexpected end of the payload inside an index list")
        self.pos += used
        return out

    def u16_varint_vec(self) -> np.ndarray:
        n = self.varint()
        if n > len(self.buf) - self.pos:      # every element takes at least one byte
            raise SerialiseError("DecodeError", "unexpected end of the payload inside a bf16 vector")
        v = self.usize_vec_n(n)
        if v.size and int(v.max()) > 0xFFFF:
            raise SerialiseError("DecodeError", "bf16 bit pattern out of range")
        return v.astype(np.uint16)

    def usize_vec_n(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint64)
        f = lib().annb_varint_decode_u64
        f.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p]
        f.restype = C.c_int64
        view = self.buf[self.pos:]
        used = f(view, len(view), n, C.c_void_p(out.ctypes.data))
        if used < 0:
            raise SerialiseError("DecodeError", "unexpected end of the payload inside an integer vector")
        self.pos += used
        return out

    def i8_vec(self) -> np.ndarray:
        n = self.varint()
        self._need(n)
        a = np.frombuffer(self.buf, dtype=np.int8, count=n, offset=self.pos).copy()
        self.pos += n
        return a

    def i32_zigzag_vec(self) -> np.ndarray:
        n = self.varint()
        if n > len(self.buf) - self.pos:
            raise SerialiseError("DecodeError", "unexpected end of the payload inside an i32 vector")
        z = self.usize_vec_n(n)
        return ((z >> np.uint64(1)).astype(np.int64) ^ -(z & np.uint64(1)).astype(np.int64)).astype(np.int32)

    def dist(self) -> int:
        v = self.varint()
        if v >= len(DIST_NAMES):
            raise SerialiseError("DecodeError", f"unknown Dist variant {v}")
        return v


def _u16_varint_vec(a) -> bytes:
    """Vec<bf16> (u16 bit patterns): length, then one variable-length integer per element."""
    a = np.ascontiguousarray(a, dtype=np.uint16).reshape(-1)
    small = a < 251
    out = np.empty(a.size * 3, dtype=np.uint8)
    # positions: 1 byte for small values, 3 for the rest
    width = np.where(small, 1, 3).astype(np.int64)
    pos = np.concatenate([[0], np.cumsum(width)[:-1]]) if a.size else np.zeros(0, np.int64)
    out[pos[small]] = a[small].astype(np.uint8)
    big = ~small
    out[pos[big]] = 251
    out[pos[big] + 1] = (a[big] & 0xFF).astype(np.uint8)
    out[pos[big] + 2] = (a[big] >> 8).astype(np.uint8)
    total = int(width.sum()) if a.size else 0
    return _varint(a.size) + out[:total].tobytes()


def _i8_vec(a) -> bytes:
    a = np.ascontiguousarray(a, dtype=np.int8).reshape(-1)
    return _varint(a.size) + a.tobytes()


def _i32_zigzag_vec(a) -> bytes:
    a = np.ascontiguousarray(a, dtype=np.int64).reshape(-1)
    z = ((a << 1) ^ (a >> 63)).astype(np.uint64)          # zig-zag of an i32 widened to 64 bits
    return encode_usize_vec(z)


def _f32_vec(a) -> bytes:
    a = np.ascontiguousarray(a, dtype="<f4").reshape(-1)
    return _varint(a.size) + a.tobytes()


def _header(kind: str, float_width: int) -> bytes:
    k = kind.encode()
    if len(k) > 255:   # one byte holds the tag length (src/serialise/mod.rs:84-92, test :1526-1535)
        raise SerialiseError("EncodeError", f"index kind tag '{kind}' is {len(k)} bytes; the header allows 255")
    return MAGIC + FORMAT_VERSION.to_bytes(4, "little") + bytes([float_width, len(k)]) + k


def _read_header(buf: bytes, path: str, kind: str, float_width: int) -> int:
    """read_header (src/serialise/mod.rs:108-170): magic, version, kind, float width, in that order."""
    if len(buf) < 8 or buf[:8] != MAGIC:
        raise SerialiseError("NotAnIndexFile", path, path=path)
    if len(buf) < 12:
        raise SerialiseError("TruncatedIndexFile", path, path=path)
    version = int.from_bytes(buf[8:12], "little")
    if version != FORMAT_VERSION:
        raise SerialiseError("UnsupportedFormatVersion", f"found {version}, supported {FORMAT_VERSION}", found=version, supported=FORMAT_VERSION)
    if len(buf) < 14:
        raise SerialiseError("TruncatedIndexFile", path, path=path)
    width, klen = buf[12], buf[13]
    if len(buf) < 14 + klen:
        raise SerialiseError("TruncatedIndexFile", path, path=path)
    found = buf[14:14 + klen].decode("utf-8", "replace")
    if found != kind:
        raise SerialiseError("IndexKindMismatch", f"expected '{kind}', found '{found}'", expected=kind, found=found)
    if width != float_width:
        raise SerialiseError("FloatWidthMismatch", f"expected {float_width}, found {width}", expected=float_width, found=width)
    return 14 + klen


def _write(dir_path: str, payload: bytes):
    """save_index (src/serialise/mod.rs:256-296): written under a temporary name and renamed into place."""
    os.makedirs(dir_path, exist_ok=True)
    final = os.path.join(dir_path, INDEX_FILE)
    tmp = final + ".tmp"
    with open(tmp, "wb") as fh:
        fh.write(payload)
    os.replace(tmp, final)


def _read(dir_path: str):
    path = os.path.join(dir_path, INDEX_FILE)
    try:
        with open(path, "rb") as fh:
            return fh.read(), path
    except OSError as e:
        raise SerialiseError("IoError", str(e)) from e


def _finish(r: _Reader, path: str):
    if r.pos != len(r.buf):   # bincode stops at the end of the payload; anything after it is an error (:316-322)
        raise SerialiseError("TrailingBytes", path, path=path)


# ----------------------------------------------------------------------------------------------- kinds
def save_exhaustive(dir_path: str, vectors, metric: int, norms=None):
    """ExhaustiveIndex<f32>::save_index.  `metric`: 0 SquaredEuclidean, 1 Cosine (annb200.L2 / annb200.COSINE);
    `norms`: the per-row L2 norms the cosine index stores (empty for Euclidean, exhaustive.rs:86-96)."""
    v = np.ascontiguousarray(vectors, dtype=np.float32)
    n, dim = v.shape
    nr = np.zeros(0, np.float32) if norms is None else np.asarray(norms, dtype=np.float32)
    _write(dir_path, _header("exhaustive", 4) + _f32_vec(v) + _varint(dim) + _varint(n) + _f32_vec(nr) + _varint(metric))


def load_exhaustive(dir_path: str) -> dict:
    buf, path = _read(dir_path)
    r = _Reader(buf, _read_header(buf, path, "exhaustive", 4))
    flat = r.f32_vec()
    dim, n = r.varint(), r.varint()
    norms = r.f32_vec()
    metric = r.dist()
    _finish(r, path)
    if flat.size != n * dim:
        raise SerialiseError("DecodeError", f"{flat.size} vector elements for n = {n}, dim = {dim}")
    return {"vectors": flat.reshape(n, dim), "dim": dim, "n": n, "norms": norms, "metric": metric}


def save_ivf(dir_path: str, vectors, metric: int, centroids, offsets, original_ids, norms=None, centroid_norms=None, all_indices=None):
    """IvfIndex<f32>::save_index.  `vectors` are in list order (after optimise_memory_layout, ivf.rs:257-294, which also
    empties `all_indices`); `offsets` [nlist + 1]; `original_ids` [n] list position -> original row."""
    v = np.ascontiguousarray(vectors, dtype=np.float32)
    n, dim = v.shape
    c = np.ascontiguousarray(centroids, dtype=np.float32)
    empty = np.zeros(0, np.float32)
    body = (_f32_vec(v) + _varint(dim) + _varint(n) + _f32_vec(empty if norms is None else norms) + _varint(metric) +
            _f32_vec(c) + _f32_vec(empty if centroid_norms is None else centroid_norms) +
            encode_usize_vec(np.zeros(0, np.uint64) if all_indices is None else all_indices) + encode_usize_vec(offsets) +
            _varint(c.shape[0]) + encode_usize_vec(original_ids))
    _write(dir_path, _header("ivf", 4) + body)


def load_ivf(dir_path: str) -> dict:
    buf, path = _read(dir_path)
    r = _Reader(buf, _read_header(buf, path, "ivf", 4))
    flat = r.f32_vec()
    dim, n = r.varint(), r.varint()
    norms = r.f32_vec()
    metric = r.dist()
    cent = r.f32_vec()
    cent_norms = r.f32_vec()
    all_indices, offsets = r.usize_vec(), r.usize_vec()
    nlist = r.varint()
    original_ids = r.usize_vec()
    _finish(r, path)
    d = {"dim": dim, "n": n, "norms": norms, "metric": metric, "centroids": cent, "centroid_norms": cent_norms, "all_indices": all_indices,
         "offsets": offsets, "nlist": nlist, "original_ids": original_ids}
    _check_ivf_sizes(d, flat.size)
    d["vectors"], d["centroids"] = flat.reshape(n, dim), cent.reshape(nlist, dim)
    return d


def _check_ivf_sizes(d: dict, n_vec_elems: int):
    """Every size the constructors rely on (a truncated or corrupt file must fail here, not in a device copy)."""
    n, dim, nlist = d["n"], d["dim"], d["nlist"]
    ok = (n_vec_elems == n * dim and d["centroids"].size == nlist * dim and d["offsets"].size == nlist + 1 and d["original_ids"].size == n and
          d["norms"].size in (0, n) and d["centroid_norms"].size in (0, nlist) and (d["offsets"].size == 0 or int(d["offsets"][-1]) == n))
    if not ok:
        raise SerialiseError("DecodeError", "inconsistent sizes in the ivf payload")


def save_exhaustive_bf16(dir_path: str, vectors_bf16, metric: int, norms=None):
    """ExhaustiveIndexBf16<f32>::save_index; `vectors_bf16` [n, dim] uint16 bit patterns, `norms` f32 norms of the un-rounded rows."""
    v = np.ascontiguousarray(vectors_bf16, dtype=np.uint16)
    n, dim = v.shape
    nr = np.zeros(0, np.float32) if norms is None else np.asarray(norms, dtype=np.float32)
    _write(dir_path, _header("exhaustive_bf16", 4) + _u16_varint_vec(v) + _varint(dim) + _varint(n) + _f32_vec(nr) + _varint(metric))


def load_exhaustive_bf16(dir_path: str) -> dict:
    buf, path = _read(dir_path)
    r = _Reader(buf, _read_header(buf, path, "exhaustive_bf16", 4))
    flat = r.u16_varint_vec()
    dim, n = r.varint(), r.varint()
    norms = r.f32_vec()
    metric = r.dist()
    _finish(r, path)
    if flat.size != n * dim or norms.size not in (0, n):
        raise SerialiseError("DecodeError", "inconsistent sizes in the exhaustive_bf16 payload")
    return {"vectors": flat.reshape(n, dim), "dim": dim, "n": n, "norms": norms, "metric": metric}


def save_exhaustive_sq8(dir_path: str, codes, metric: int, scales, norms_i=None):
    """ExhaustiveSq8Index<f32>::save_index; `codes` [n, dim] int8, `norms_i` the i32 code norms (cosine), `scales` the codebook."""
    v = np.ascontiguousarray(codes, dtype=np.int8)
    n, dim = v.shape
    ni = np.zeros(0, np.int32) if norms_i is None else np.asarray(norms_i, dtype=np.int32)
    _write(dir_path, _header("exhaustive_sq8", 4) + _i8_vec(v) + _i32_zigzag_vec(ni) + _varint(dim) + _varint(n) + _varint(metric) + _f32_vec(scales))


def load_exhaustive_sq8(dir_path: str) -> dict:
    buf, path = _read(dir_path)
    r = _Reader(buf, _read_header(buf, path, "exhaustive_sq8", 4))
    codes = r.i8_vec()
    norms_i = r.i32_zigzag_vec()
    dim, n = r.varint(), r.varint()
    metric = r.dist()
    scales = r.f32_vec()
    _finish(r, path)
    if codes.size != n * dim or norms_i.size not in (0, n) or scales.size != dim:
        raise SerialiseError("DecodeError", "inconsistent sizes in the exhaustive_sq8 payload")
    return {"vectors": codes.reshape(n, dim), "dim": dim, "n": n, "norms_i": norms_i, "metric": metric, "scales": scales}


def save_ivf_bf16(dir_path: str, vectors_bf16, metric: int, centroids, offsets, original_ids, norms=None, centroid_norms=None, all_indices=None):
    """IvfIndexBf16<f32>::save_index (list order, as `build` leaves it)."""
    v = np.ascontiguousarray(vectors_bf16, dtype=np.uint16)
    n, dim = v.shape
    c = np.ascontiguousarray(centroids, dtype=np.float32)
    empty = np.zeros(0, np.float32)
    body = (_u16_varint_vec(v) + _varint(dim) + _varint(n) + _f32_vec(empty if norms is None else norms) + _varint(metric) +
            _f32_vec(c) + _f32_vec(empty if centroid_norms is None else centroid_norms) +
            encode_usize_vec(np.zeros(0, np.uint64) if all_indices is None else all_indices) + encode_usize_vec(offsets) +
            _varint(c.shape[0]) + encode_usize_vec(original_ids))
    _write(dir_path, _header("ivf_bf16", 4) + body)


def load_ivf_bf16(dir_path: str) -> dict:
    buf, path = _read(dir_path)
    r = _Reader(buf, _read_header(buf, path, "ivf_bf16", 4))
    flat = r.u16_varint_vec()
    dim, n = r.varint(), r.varint()
    norms = r.f32_vec()
    metric = r.dist()
    cent = r.f32_vec()
    cent_norms = r.f32_vec()
    all_indices, offsets = r.usize_vec(), r.usize_vec()
    nlist = r.varint()
    original_ids = r.usize_vec()
    _finish(r, path)
    d = {"dim": dim, "n": n, "norms": norms, "metric": metric, "centroids": cent, "centroid_norms": cent_norms, "all_indices": all_indices,
         "offsets": offsets, "nlist": nlist, "original_ids": original_ids}
    _check_ivf_sizes(d, flat.size)
    d["vectors"], d["centroids"] = flat.reshape(n, dim), cent.reshape(nlist, dim)
    return d


def save_ivf_sq8(dir_path: str, codes, metric: int, centroids, offsets, original_ids, scales, norms_i=None, centroid_norms=None, all_indices=None):
    """IvfSq8Index<f32>::save_index (list order)."""
    v = np.ascontiguousarray(codes, dtype=np.int8)
    n, dim = v.shape
    c = np.ascontiguousarray(centroids, dtype=np.float32)
    ni = np.zeros(0, np.int32) if norms_i is None else np.asarray(norms_i, dtype=np.int32)
    body = (_i8_vec(v) + _i32_zigzag_vec(ni) + _varint(dim) + _varint(n) + _varint(metric) + _f32_vec(c) +
            _f32_vec(np.zeros(0, np.float32) if centroid_norms is None else centroid_norms) +
            encode_usize_vec(np.zeros(0, np.uint64) if all_indices is None else all_indices) + encode_usize_vec(offsets) + _f32_vec(scales) +
            _varint(c.shape[0]) + encode_usize_vec(original_ids))
    _write(dir_path, _header("ivf_sq8", 4) + body)


def load_ivf_sq8(dir_path: str) -> dict:
    buf, path = _read(dir_path)
    r = _Reader(buf, _read_header(buf, path, "ivf_sq8", 4))
    codes = r.i8_vec()
    norms_i = r.i32_zigzag_vec()
    dim, n = r.varint(), r.varint()
    metric = r.dist()
    cent = r.f32_vec()
    cent_norms = r.f32_vec()
    all_indices, offsets = r.usize_vec(), r.usize_vec()
    scales = r.f32_vec()
    nlist = r.varint()
    original_ids = r.usize_vec()
    _finish(r, path)
    d = {"dim": dim, "n": n, "norms": np.zeros(0, np.float32), "norms_i": norms_i, "metric": metric, "centroids": cent, "centroid_norms": cent_norms,
         "all_indices": all_indices, "offsets": offsets, "scales": scales, "nlist": nlist, "original_ids": original_ids}
    _check_ivf_sizes(d, codes.size)
    if norms_i.size not in (0, n) or scales.size != dim:
        raise SerialiseError("DecodeError", "inconsistent sizes in the ivf_sq8 payload")
    d["vectors"], d["centroids"] = codes.reshape(n, dim), cent.reshape(nlist, dim)
    return d


# ----------------------------------------------------------------------------------------------- serving a saved index
def load_exhaustive_b200(dir_path: str, device: int = 0):
    """A saved CPU `ExhaustiveIndex<f32>` as a resident B200 index (the norms are recomputed on the device in the same order)."""
    from . import ExhaustiveIndexB200, F32
    d = load_exhaustive(dir_path)
    if d["metric"] == 2:
        raise SerialiseError("DecodeError", "Manhattan indices have no B200 counterpart (src/gpu/exhaustive_gpu.rs:73-75)")
    return ExhaustiveIndexB200.new(d["vectors"], d["metric"], F32, device=device)


def load_ivf_b200(dir_path: str, device: int = 0):
    """A saved CPU `IvfIndex<f32>` (list-ordered, as `build` leaves it) as a resident B200 index."""
    from . import IvfIndexB200, F32
    d = load_ivf(dir_path)
    if d["metric"] == 2:
        raise SerialiseError("DecodeError", "Manhattan indices have no B200 counterpart (src/cpu/ivf.rs:153-155)")
    if d["all_indices"].size:
        raise SerialiseError("DecodeError", "index was saved before optimise_memory_layout: vectors are not in list order")
    cos = d["metric"] == 1
    return IvfIndexB200.from_parts(d["vectors"], d["centroids"], d["offsets"], d["original_ids"], F32, d["metric"],
                                   norms=d["norms"] if cos else None, centroid_norms=d["centroid_norms"] if cos else None, device=device)


def load_ivf_quantised_b200(dir_path: str, kind: str, device=0):
    """A saved `IvfIndexBf16<f32>` ("ivf_bf16") or `IvfSq8Index<f32>` ("ivf_sq8") as a resident B200 index."""
    from . import BF16, SQ8, IvfIndexB200
    d = load_ivf_bf16(dir_path) if kind == "ivf_bf16" else load_ivf_sq8(dir_path)
    if d["metric"] == 2:
        raise SerialiseError("DecodeError", "Manhattan indices have no B200 counterpart")
    if d["all_indices"].size:
        raise SerialiseError("DecodeError", "index was saved before optimise_memory_layout: vectors are not in list order")
    cos = d["metric"] == 1
    if kind == "ivf_bf16":
        return IvfIndexB200.from_parts(d["vectors"], d["centroids"], d["offsets"], d["original_ids"], BF16, d["metric"],
                                       norms=d["norms"] if cos else None, centroid_norms=d["centroid_norms"] if cos else None, device=device)
    return IvfIndexB200.from_parts(d["vectors"], d["centroids"], d["offsets"], d["original_ids"], SQ8, d["metric"],
                                   norms=d["norms_i"] if cos else None, sq8_scales=d["scales"], device=device)
