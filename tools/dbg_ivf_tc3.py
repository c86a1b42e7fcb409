import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python"), os.path.join(ROOT, "tests")]
import numpy as np
import annb200
from oracle import oracle as o
n, dim = 640, 32
rng = np.random.default_rng(1)
data = rng.standard_normal((n, dim)).astype(np.float32)
c = o.build_ivf(data, o.L2, nlist=1, centroids=np.zeros((1, dim), np.float32))
g = annb200.IvfIndexB200.from_parts(c.vectors, c.centroids, c.offsets, c.original_ids, c.dtype, c.metric, norms=c.norms, centroid_norms=c.centroid_norms, sq8_scales=c.scales, list_begin=0, list_end=1, n_total=n)
g.set_option("ivf_list_major", 1); g.set_option("path", annb200.PATH_TENSOR); g.set_option("cert_fallback", 0); g.set_option("tc_candidates", 32)
x64 = c.vectors.astype(np.float64)
def run(name, q):
    q = np.ascontiguousarray(q[None, :], np.float32)
    g.query_batch(q, 24, nprobe=1)
    raw = g.debug_fetch_tile().view(np.uint64).ravel()
    want = (x64 * x64).sum(1) - 2 * x64 @ q[0].astype(np.float64)
    errs = []
    for half in (0, 1):
        keys = raw[half * 32:(half + 1) * 32]
        idx = (keys & 0xFFFFFFFF).astype(np.int64)
        ob = (keys >> 32).astype(np.uint32)
        val = np.where(ob & 0x80000000, ob ^ 0x80000000, ~ob).astype(np.uint32).view(np.float32)
        errs += [(int(i), round(float(v - want[i]), 3)) for i, v in zip(idx, val) if 0 <= i < n]
    print(name, "max |err|", max(abs(e) for _, e in errs), errs[:10])
run("zero", np.zeros(dim))
for j in (0, 1, 7, 8, 15, 16, 31):
    e = np.zeros(dim); e[j] = 1.0
    run(f"e{j}", e)
run("ones", np.ones(dim))
run("row0", data[0])
