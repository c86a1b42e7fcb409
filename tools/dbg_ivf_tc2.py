"""Debug helper: one list, a few queries, k' = 32 on the IVF tensor path -- which list positions are found?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python"), os.path.join(ROOT, "tests")]
import numpy as np
import annb200
from oracle import datagen, oracle as o
n, dim = 640, 32
rng = np.random.default_rng(1)
data = rng.standard_normal((n, dim)).astype(np.float32)
cent = np.zeros((1, dim), np.float32)
c = o.build_ivf(data, o.L2, nlist=1, centroids=cent)
g = annb200.IvfIndexB200.from_parts(c.vectors, c.centroids, c.offsets, c.original_ids, c.dtype, c.metric, norms=c.norms, centroid_norms=c.centroid_norms, sq8_scales=c.scales, list_begin=0, list_end=1, n_total=n)
g.set_option("ivf_list_major", 1)
g.set_option("path", annb200.PATH_TENSOR)
g.set_option("cert_fallback", 0)
for cand in (16, 32):
    g.set_option("tc_candidates", cand)
    for nq in (1, 3, 130):
        q = data[:nq] + 0.001
        k = 24 if cand == 32 else 10
        got = g.query_batch(q, k, nprobe=1)
        ref = o.ivf_search(c, q, k, nprobe=1)
        pos_got = sorted(int(np.where(c.original_ids == i)[0][0]) for i in got[0][0] if i >= 0)
        pos_ref = sorted(int(np.where(c.original_ids == i)[0][0]) for i in ref[0][0])
        print("cand", cand, "nq", nq, "rows_ok", int((got[0] == ref[0]).all(1).sum()), "/", nq)
        print("  got pos", pos_got)
        print("  ref pos", pos_ref)
        if nq == 1:
            raw = g.debug_fetch_tile().view(np.uint64).ravel()
            x64 = c.vectors.astype(np.float64); q64 = q[0].astype(np.float64)
            want = (x64 * x64).sum(1) - 2 * x64 @ q64
            for half in (0, 1):
                keys = raw[half * cand:(half + 1) * cand]
                idx = (keys & 0xFFFFFFFF).astype(np.int64)
                ob = (keys >> 32).astype(np.uint32)
                fb = np.where(ob & 0x80000000, ob ^ 0x80000000, ~ob).astype(np.uint32)
                val = fb.view(np.float32)
                print("  half", half, [(int(i), round(float(v), 3), round(float(want[i]), 3) if 0 <= i < n else None) for i, v in zip(idx[:12], val[:12])])
            order = np.argsort(want)
            print("  true smallest", [(int(i), round(float(want[i]), 3)) for i in order[:12]])
