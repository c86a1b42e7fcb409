"""Build-side timing: annb_kmeans_lloyd (250k x 128 training rows, 4096 centroids) and annb_ivf_assign (2M x 128 rows) on the
tensor-core assignment path and, with ANNB200_ASSIGN_PATH=simt, on the exact CUDA-core kernel; checks that both agree."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + "/ann-search-rs_b200/python"]
import numpy as np, annb200
from annb200 import datagen
redone = annb200.lib().annb_assign_last_redone
redone.restype = C.c_uint64
d = datagen.correlated(250_000, 128, seed=42)
init = d[(np.arange(4096) * 250_000) // 4096].copy()
annb200.kmeans_lloyd(d[:10000], init[:64], annb200.L2, 1)
res = {}
for path in ("tensor", "simt"):
    os.environ["ANNB200_ASSIGN_PATH"] = path
    t = time.perf_counter(); c, it = annb200.kmeans_lloyd(d, init, annb200.L2, 8); dt = time.perf_counter() - t
    res[path] = c
    print(f"[{path}] device Lloyd 250k x 128, nlist 4096: {it} updates in {dt:.2f} s ({dt / max(it, 1) * 1e3:.0f} ms per iteration incl. upload), rows redone exactly {int(redone())}", flush=True)
print("Lloyd centroids identical:", bool(np.array_equal(res["tensor"].view(np.uint32), res["simt"].view(np.uint32))))
big = np.concatenate([d] * 8)   # 2M rows
a = {}
for path in ("tensor", "simt"):
    os.environ["ANNB200_ASSIGN_PATH"] = path
    annb200.ivf_assign(big[:100000], res["simt"], annb200.L2)
    t = time.perf_counter(); a[path] = annb200.ivf_assign(big, res["simt"], annb200.L2); dt = time.perf_counter() - t
    print(f"[{path}] annb_ivf_assign 2M x 128 -> 4096 cells: {dt * 1e3:.0f} ms incl. host->device copy of the rows, rows redone exactly {int(redone())}", flush=True)
print("assignments identical:", bool(np.array_equal(a["tensor"], a["simt"])))
