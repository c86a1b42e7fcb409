"""Host-side reader / writer of the crate's saved indices (annb200.serialise; src/serialise/mod.rs).  The header and
error behaviour restate the reference's own tests (src/serialise/mod.rs:1514-1640); the payload is checked by round trips and
against bincode's specified integer encoding (its bytes are unpinned: no saved index ships with the reference)."""
import os

import numpy as np
import pytest

import annb200
from annb200 import serialise as S
from annb200 import datagen
from oracle import oracle as o


def _saved_exhaustive(tmp_path, metric=0):
    data = datagen.gaussian_noise(64, 16, seed=7)
    norms = np.array([o.l2_norm_f32(r) for r in data], np.float32) if metric == 1 else None
    S.save_exhaustive(str(tmp_path), data, metric, norms)
    return data, norms, os.path.join(str(tmp_path), S.INDEX_FILE)


def test_varint_boundaries_follow_the_bincode_specification():
    cases = {0: "00", 250: "fa", 251: "fbfb00", 65535: "fbffff", 65536: "fc00000100", 2**32 - 1: "fcffffffff",
             2**32: "fd0000000001000000", 2**64 - 1: "fd" + "ff" * 8}
    for v, hx in cases.items():
        assert S._varint(v).hex() == hx, v
    vals = np.array(list(cases), dtype=np.uint64)
    enc = S.encode_usize_vec(vals)
    assert enc.hex() == "08" + "".join(cases.values())
    r = S._Reader(enc, 0)
    assert np.array_equal(r.usize_vec(), vals) and r.pos == len(enc)
    rng = np.random.default_rng(1)
    big = np.concatenate([rng.integers(0, 300, 5000), rng.integers(0, 1 << 20, 5000), rng.integers(0, 1 << 40, 5000)]).astype(np.uint64)
    enc = S.encode_usize_vec(big)
    assert enc == S._varint(big.size) + b"".join(S._varint(int(v)) for v in big)
    assert np.array_equal(S._Reader(enc, 0).usize_vec(), big)


def test_exhaustive_bytes_of_a_tiny_index(tmp_path):
    S.save_exhaustive(str(tmp_path), np.array([[1.0, 2.0]], np.float32), 0)
    raw = open(os.path.join(str(tmp_path), S.INDEX_FILE), "rb").read()
    want = (b"ANNSRS\0\0" + (2).to_bytes(4, "little") + bytes([4, 10]) + b"exhaustive" +      # header (mod.rs:84-106)
            bytes([2]) + np.array([1.0, 2.0], "<f4").tobytes() + bytes([2, 1, 0, 0]))          # vectors_flat, dim, n, norms (empty), metric
    assert raw == want


@pytest.mark.parametrize("metric", [0, 1])
def test_exhaustive_round_trip(tmp_path, metric):
    data, norms, _ = _saved_exhaustive(tmp_path, metric)
    d = S.load_exhaustive(str(tmp_path))
    assert d["n"] == 64 and d["dim"] == 16 and d["metric"] == metric
    assert np.array_equal(d["vectors"].view(np.uint32), data.view(np.uint32))
    assert d["norms"].size == (64 if metric else 0)
    if metric:
        assert np.array_equal(d["norms"].view(np.uint32), norms.view(np.uint32))


@pytest.mark.parametrize("metric", [o.L2, o.COSINE])
def test_ivf_round_trip(tmp_path, metric):
    data = datagen.gaussian_noise(3000, 12, seed=9)
    ix = o.build_ivf(data, metric, nlist=40, kmeans_iters=3)
    S.save_ivf(str(tmp_path), ix.vectors, int(metric == o.COSINE), ix.centroids, ix.offsets, ix.original_ids, norms=ix.norms, centroid_norms=ix.centroid_norms)
    d = S.load_ivf(str(tmp_path))
    assert (d["n"], d["dim"], d["nlist"]) == (3000, 12, 40) and d["all_indices"].size == 0
    assert np.array_equal(d["vectors"].view(np.uint32), np.asarray(ix.vectors, np.float32).view(np.uint32))
    assert np.array_equal(d["centroids"], ix.centroids) and np.array_equal(d["offsets"], np.asarray(ix.offsets, np.uint64))
    assert np.array_equal(d["original_ids"], np.asarray(ix.original_ids, np.uint64))
    if metric == o.COSINE:
        assert np.array_equal(d["norms"], ix.norms) and np.array_equal(d["centroid_norms"], ix.centroid_norms)
    with pytest.raises(S.SerialiseError) as e:          # mod.rs:1616-1625
        S.load_exhaustive(str(tmp_path))
    assert e.value.variant == "IndexKindMismatch" and e.value.fields == {"expected": "exhaustive", "found": "ivf"}


def test_rejects_files_that_are_not_indices_or_are_cut_short(tmp_path):
    p = os.path.join(str(tmp_path), S.INDEX_FILE)
    open(p, "wb").write(b"definitely not an index")                    # mod.rs:1514-1521
    with pytest.raises(S.SerialiseError) as e:
        S.load_exhaustive(str(tmp_path))
    assert e.value.variant == "NotAnIndexFile"
    open(p, "wb").write(S.MAGIC[:4])                                    # mod.rs:1537-1545
    with pytest.raises(S.SerialiseError) as e:
        S.load_exhaustive(str(tmp_path))
    assert e.value.variant == "NotAnIndexFile"
    _, _, path = _saved_exhaustive(tmp_path)
    raw = open(path, "rb").read()
    for cut in (8, 12, 14):                                             # mod.rs:1548-1564
        open(path, "wb").write(raw[:cut])
        with pytest.raises(S.SerialiseError) as e:
            S.load_exhaustive(str(tmp_path))
        assert e.value.variant == "TruncatedIndexFile", cut
    open(path, "wb").write(raw[:len(raw) // 2])                          # mod.rs:1567-1578
    with pytest.raises(S.SerialiseError) as e:
        S.load_exhaustive(str(tmp_path))
    assert e.value.variant == "DecodeError"
    open(path, "wb").write(raw + bytes(28))                              # mod.rs:1581-1594
    with pytest.raises(S.SerialiseError) as e:
        S.load_exhaustive(str(tmp_path))
    assert e.value.variant == "TrailingBytes"
    bad = bytearray(raw)
    bad[8:12] = (S.FORMAT_VERSION + 1).to_bytes(4, "little")              # mod.rs:1597-1613
    open(path, "wb").write(bytes(bad))
    with pytest.raises(S.SerialiseError) as e:
        S.load_exhaustive(str(tmp_path))
    assert e.value.variant == "UnsupportedFormatVersion" and e.value.fields == {"found": 3, "supported": 2}
    bad = bytearray(raw)
    bad[12] = 8                                                          # an f64 index: mod.rs:1627-1638
    open(path, "wb").write(bytes(bad))
    with pytest.raises(S.SerialiseError) as e:
        S.load_exhaustive(str(tmp_path))
    assert e.value.variant == "FloatWidthMismatch" and e.value.fields == {"expected": 4, "found": 8}
    with pytest.raises(S.SerialiseError) as e:                           # mod.rs:1641-1647
        S.load_exhaustive("/nonexistent/ann-search-rs")
    assert e.value.variant == "IoError"
    with pytest.raises(S.SerialiseError) as e:                           # mod.rs:1526-1535
        S._header("a" * 256, 4)
    assert e.value.variant == "EncodeError"
    assert len(S._header("a" * 255, 4)) == 14 + 255


def _build_cpp_tool():
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "ann-search-rs_b200")
    os.makedirs(os.path.join(pkg, "build"), exist_ok=True)
    exe = os.path.join(pkg, "build", "serialise_roundtrip")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", exe, os.path.join(pkg, "host", "tests", "serialise_roundtrip.cpp"),
           "-L" + os.path.join(pkg, "lib"), "-lannb200", "-Wl,-rpath," + os.path.join(pkg, "lib"),
           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_mirror_reads_and_rewrites_the_same_bytes(tmp_path):
    """host/annb200_serialise.hpp and annb200/serialise.py agree byte for byte, and name the same error variants."""
    import subprocess
    exe = _build_cpp_tool()
    a, b = tmp_path / "py", tmp_path / "cpp"
    for d in (a / "ex", a / "ivf", b / "ex", b / "ivf"):
        d.mkdir(parents=True)
    data = datagen.gaussian_noise(500, 10, seed=3)
    norms = np.array([o.l2_norm_f32(r) for r in data], np.float32)
    S.save_exhaustive(str(a / "ex"), data, 1, norms)
    ix = o.build_ivf(data, o.COSINE, nlist=300, kmeans_iters=2)          # 300 lists: offsets cross the one-byte varint range
    S.save_ivf(str(a / "ivf"), ix.vectors, 1, ix.centroids, ix.offsets, ix.original_ids, norms=ix.norms, centroid_norms=ix.centroid_norms)
    for kind, sub in (("exhaustive", "ex"), ("ivf", "ivf")):
        r = subprocess.run([exe, kind, str(a / sub), str(b / sub)], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.strip() == "OK", r.stdout + r.stderr
        assert open(a / sub / S.INDEX_FILE, "rb").read() == open(b / sub / S.INDEX_FILE, "rb").read(), kind
    raw = open(a / "ex" / S.INDEX_FILE, "rb").read()
    cases = {"NotAnIndexFile": b"definitely not an index", "TruncatedIndexFile": raw[:12], "DecodeError": raw[:len(raw) // 2],
             "TrailingBytes": raw + bytes(28), "UnsupportedFormatVersion": raw[:8] + (3).to_bytes(4, "little") + raw[12:],
             "FloatWidthMismatch": raw[:12] + bytes([8]) + raw[13:]}
    for want, content in cases.items():
        open(b / "ex" / S.INDEX_FILE, "wb").write(content)
        r = subprocess.run([exe, "variant", "exhaustive", str(b / "ex")], capture_output=True, text=True)
        assert r.stdout.strip() == want, (want, r.stdout)
    r = subprocess.run([exe, "variant", "exhaustive", str(a / "ivf")], capture_output=True, text=True)
    assert r.stdout.strip() == "IndexKindMismatch"
    r = subprocess.run([exe, "variant", "ivf", str(tmp_path / "missing")], capture_output=True, text=True)
    assert r.stdout.strip() == "IoError"
