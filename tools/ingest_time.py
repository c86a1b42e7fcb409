"""matrix_to_flat timing: a column-major (faer-layout) host matrix -> row-major, on the device (annb_matrix_to_flat, host result)
against numpy's single-threaded strided copy on the host (the shape of the reference's src/utils/mod.rs:44-68 loop)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + "/ann-search-rs_b200/python"]
import numpy as np, annb200
n, dim = (int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000), 128
a = np.asfortranarray(np.random.default_rng(0).standard_normal((n, dim), dtype=np.float32))
annb200.matrix_to_flat(a[:1000])
t = time.perf_counter(); g = annb200.matrix_to_flat(a); tg = time.perf_counter() - t
t = time.perf_counter(); h = np.ascontiguousarray(a); th = time.perf_counter() - t
print(f"matrix_to_flat {n} x {dim} column-major -> row-major: device path {tg * 1e3:.0f} ms (pageable host in, host out), numpy on one core {th * 1e3:.0f} ms, equal {bool(np.array_equal(g, h))}")
