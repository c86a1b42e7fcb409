"""Diagnostic: error of the first tile's selection values against a float64 GEMM, for several row widths / operand placements."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python")]
import annb200
from annb200 import datagen

def tf32(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return u.view(np.float32)

for dim, opts in [(128, {}), (128, {"tc_f32_lo_smem": 1}), (256, {}), (192, {}), (64, {}), (64, {"tc_f32_lo_smem": 1})]:
    data = datagen.gaussian_noise(8192, dim, seed=3)
    q = datagen.subsample_with_noise(data, 128, seed=3)
    g = annb200.ExhaustiveIndexB200.new(data, annb200.L2, annb200.F32)
    g.set_option("path", annb200.PATH_TENSOR)
    g.set_option("tc_debug", 1); g.set_option("db_splits", 1)
    for k_, v_ in opts.items(): g.set_option(k_, v_)
    g.query_batch(q, 10)
    v = g.debug_fetch_tile().astype(np.float64)
    x = data[:128].astype(np.float64)
    want = (x * x).sum(1)[None, :] - 2 * (q.astype(np.float64) @ x.T)
    scale = np.abs(want).max()
    qh = tf32(q); ql = tf32(q - qh); xh = tf32(data[:128]); xl = tf32(data[:128] - xh)
    def val(terms):
        s = sum(a.astype(np.float64) @ b.astype(np.float64).T for a, b in terms)
        return (x * x).sum(1)[None, :] - 2 * s
    e3 = np.abs(val([(qh, xh), (ql, xh), (qh, xl)]) - want).max() / scale
    e_nolo = np.abs(val([(qh, xh), (qh, xl)]) - want).max() / scale
    e_noxl = np.abs(val([(qh, xh), (ql, xh)]) - want).max() / scale
    print(f"dim {dim} {opts}: gpu err {np.abs(v - want).max() / scale:.3e} | exact 3xTF32 {e3:.3e} | without Qlo.Xhi {e_nolo:.3e} | without Qhi.Xlo {e_noxl:.3e} | gpu vs 3-term model {np.abs(v - val([(qh, xh), (ql, xh), (qh, xl)])).max() / scale:.3e} gpu vs no-Qlo model {np.abs(v - val([(qh, xh), (qh, xl)])).max() / scale:.3e}")
