"""Small end-to-end pass over every kernel family (for compute-sanitizer): flat SIMT/tensor, IVF query-major /
list-major / tensor, all dtypes, fallback path, merge.  Checks results against the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python")]
import annb200
from annb200 import datagen
from oracle import oracle as o

def same(a, b): return np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))

data = datagen.gaussian_noise(5000, 72, seed=1)
q = datagen.subsample_with_noise(data, 70, seed=1)
for dt, odt in ((annb200.F32, o.F32), (annb200.BF16, o.BF16), (annb200.SQ8, o.SQ8)):
    for met, omet in ((annb200.L2, o.L2), (annb200.COSINE, o.COSINE)):
        g = annb200.ExhaustiveIndexB200.new(data, met, dt)
        c = o.build_flat(data, omet, odt)
        ref = o.flat_search(c, q, 10)
        for path in (annb200.PATH_SIMT, annb200.PATH_AUTO):
            g.set_option("path", path)
            assert same(g.query_batch(q, 10), ref), ("flat", dt, met, path)
        if dt != annb200.SQ8:
            g.set_option("cert_eps_log2", -2)
            assert same(g.query_batch(q, 10), ref), "flat fallback"
        assert same(g.generate_knn(5, row_begin=10, row_end=60), o.flat_search(c, None, 5, self_rows=np.arange(10, 60), self_mode=True))
        g.close()
        ci = o.build_ivf(data, omet, nlist=20, dtype=odt, kmeans_iters=2)
        norms = ci.norms_i if odt == o.SQ8 else ci.norms
        gi = annb200.IvfIndexB200.from_parts(ci.vectors, ci.centroids, ci.offsets, ci.original_ids, ci.dtype, ci.metric, norms=norms,
                                             centroid_norms=ci.centroid_norms, sq8_scales=ci.scales)
        r = o.ivf_search(ci, q, 10, nprobe=5)
        for lm in (0, 1):
            for path in (annb200.PATH_SIMT, annb200.PATH_AUTO):
                gi.set_option("ivf_list_major", lm); gi.set_option("path", path)
                got = gi.query_batch(q, 10, nprobe=5)
                assert np.array_equal(got[1].view(np.uint32), r[1].view(np.uint32)), ("ivf", dt, met, lm, path)
        gi.close()
a = annb200.ivf_assign(data, data[::250].copy(), annb200.L2)
assert np.array_equal(a.astype(np.int64), o.assign_all(data, data[::250].copy(), None, o.L2))
print("sanitize smoke OK")
