"""Prints the tensor-path role wait-cycle counters of CTA (0,0) (debug instrumentation)."""
import ctypes as C, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python")]
import annb200
from annb200 import datagen
n, dim, nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 128, int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
data = datagen.correlated(n, dim, seed=42)
q = datagen.subsample_with_noise(data, nq, seed=42)
lib = annb200.lib()
lib.annb_debug_fetch_cycles.argtypes = [C.c_void_p, C.c_void_p]
for name, dt in (("f32", annb200.F32), ("bf16", annb200.BF16), ("sq8", annb200.SQ8)):
    g = annb200.ExhaustiveIndexB200.new(data, annb200.COSINE, dt)
    g.set_option("path", annb200.PATH_TENSOR)
    g.set_option("tc_debug", 1)
    for _ in range(2):
        g.query_batch(q, 10)
    out = np.zeros(8, dtype=np.uint64)
    annb200._check(lib.annb_debug_fetch_cycles(g.handle, out.ctypes.data_as(C.c_void_p)))
    tot, prod, full, tempty, tfull, slow, tiles = [int(x) for x in out[:7]]
    print(f"{name}: tiles={tiles} total={tot} cyc/tile={tot/max(tiles,1):.0f} | producer wait-empty {prod/tot:.2%} | mma wait-full {full/tot:.2%} "
          f"wait-tmem-empty {tempty/tot:.2%} | epilogue(thread 64) wait-tmem-full {tfull/tot:.2%} slow-path {slow/tot:.2%}")
    g.close()
