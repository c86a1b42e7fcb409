"""Debug helper: run the IVF tensor path on a few (dim, k) combinations, one process each, blocking launches."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    for dim, k, cand, dt in ((50, 15, 0, 0), (32, 15, 0, 0), (50, 10, 0, 0)):
        r = subprocess.run([sys.executable, __file__, str(dim), str(k), str(cand), str(dt)], capture_output=True, text=True, env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1"))
        print(dim, k, cand, dt, "rc", r.returncode, "\n".join((r.stdout + r.stderr).strip().splitlines()[-2:])[:300])
    sys.exit(0)
sys.path[:0] = [ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python"), os.path.join(ROOT, "tests")]
import numpy as np
import annb200
from oracle import datagen, oracle as o
dim, k = int(sys.argv[1]), int(sys.argv[2])
data = datagen.gaussian_noise(2500, dim, seed=21)
q = datagen.subsample_with_noise(data, 33, seed=21)
c = o.build_ivf(data, o.L2, nlist=40, dtype=int(sys.argv[4]), kmeans_iters=5)
g = annb200.IvfIndexB200.from_parts(c.vectors, c.centroids, c.offsets, c.original_ids, c.dtype, c.metric, norms=(c.norms_i if c.dtype == o.SQ8 else c.norms), centroid_norms=c.centroid_norms, sq8_scales=c.scales, list_begin=0, list_end=c.nlist, n_total=c.n)
if os.environ.get("SEQ"):
    g.query_batch(q, k, nprobe=None)
    g.set_option("ivf_list_major", 1)
    g.set_option("path", annb200.PATH_SIMT)
    g.query_batch(q, k, nprobe=None)
g.set_option("ivf_list_major", 1)
g.set_option("path", annb200.PATH_TENSOR)
if int(sys.argv[3]): g.set_option("tc_candidates", int(sys.argv[3]))
got = g.query_batch(q, k, nprobe=None)
ref = o.ivf_search(c, q, k, nprobe=None)
bad = np.argwhere(got[0] != ref[0])
print("uncert", g.get_stat("uncertified"), "fallback", g.get_stat("fallback_queries"), "mismatching cells", len(bad), "rows", len(set(bad[:, 0].tolist())))
if len(bad):
    r0 = bad[0, 0]
    print("row", r0, "got", flush=True) if False else None
    print("row", r0, "got", got[0][r0].tolist(), [float(x) for x in got[1][r0]])
    print("row", r0, "ref", ref[0][r0].tolist(), [float(x) for x in ref[1][r0]])
print("ok ids_equal", bool(np.array_equal(got[0], ref[0])), "dist_equal", bool(np.array_equal(got[1].view(np.uint32), ref[1].view(np.uint32))))
