"""Reader / writer for the crate's saved indices (`IndexIo`, src/serialise/mod.rs:33-335), kinds "exhaustive" and "ivf",
so that an index built and saved by the CPU crate can be served from a B200 and the other way round (SURVEY section 8f,
row 3).

File layout (src/serialise/mod.rs:28-106, 240-330): a directory holding `index.bin` =
    8 magic bytes b"ANNSRS\\0\\0" | u32 LE format version (2) | u8 float width | u8 kind length | kind tag |
    bincode 2 payload of the serde-derived struct, `bincode::config::standard()`: little-endian, variable-length
    integers (one byte below 251, else marker 251 / 252 / 253 + u16 / u32 / u64), f32 as 4 raw bytes, `Vec<T>` as
    length + elements, unit enum variants as their u32 index, fields in declaration order:
      ExhaustiveIndex<T> (src/cpu/exhaustive.rs:18-32): vectors_flat, dim, n, norms, metric
      IvfIndex<T>        (src/cpu/ivf.rs:24-48): vectors_flat, dim, n, norms, metric, centroids, centroids_norm,
                         all_indices, offsets, nlist, original_ids
      Dist               (src/utils/dist.rs:29-37): SquaredEuclidean = 0, Cosine = 1, Manhattan = 2

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

    def dist(self) -> int:
        v = self.varint()
        if v >= len(DIST_NAMES):
            raise SerialiseError("DecodeError", f"unknown Dist variant {v}")
        return v


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
    if flat.size != n * dim or cent.size != nlist * dim or offsets.size != nlist + 1:
        raise SerialiseError("DecodeError", "inconsistent sizes in the ivf payload")
    return {"vectors": flat.reshape(n, dim), "dim": dim, "n": n, "norms": norms, "metric": metric, "centroids": cent.reshape(nlist, dim),
            "centroid_norms": cent_norms, "all_indices": all_indices, "offsets": offsets, "nlist": nlist, "original_ids": original_ids}


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
