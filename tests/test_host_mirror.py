"""Host-side mirror (annb200 python module): the numpy build steps equal the oracle bit for bit.  CPU only."""
import numpy as np

import annb200
from oracle import oracle as o


def test_row_norms_and_normalise_match_oracle(rng):
    for dim in (3, 8, 32, 50, 128, 131):
        x = (rng.standard_normal((257, dim)) * 5).astype(np.float32)
        assert np.array_equal(annb200.ref_row_norms(x), o.row_norms_f32(x))
        assert np.array_equal(annb200.normalise_rows(x), o.normalise_rows(x))
        assert np.array_equal(annb200.seq_row_norms(x), np.array([o.seq_norm_f32(r) for r in x], np.float32))


def test_quantisers_match_oracle(rng):
    x = (rng.standard_normal((500, 37)) * 3).astype(np.float32)
    x[:, 5] = 0
    assert np.array_equal(annb200.encode_bf16(x), o.encode_bf16(x))
    sc = annb200.sq8_train(x)
    assert np.array_equal(sc, o.sq8_train(x)) and sc[5] == 1.0
    assert np.array_equal(annb200.sq8_encode(x * 2, sc), o.sq8_encode(x * 2, sc))


def test_csr_layout_matches_oracle(rng):
    a = rng.integers(0, 13, 1000)
    idx, off = annb200.build_csr_layout(a, 16)
    ridx, roff = o.build_csr(a, 16)
    assert np.array_equal(idx.astype(np.int64), ridx) and np.array_equal(off.astype(np.int64), roff)


def test_defaults():
    assert annb200.default_nlist(10_000_000) == 3162     # src/cpu/ivf.rs:172
    assert annb200.default_nprobe(4096) == 64            # src/cpu/ivf.rs:345-347
    assert annb200.default_nprobe(1) == 1


def test_bench_row_count_flag_survives_torchrun(monkeypatch):
    """torch.distributed.run abbreviates its own options, so a bare `--n` after the script name is rejected as ambiguous
    (--nnodes / --nproc-per-node / ...): multi-GPU launches pass `--rows`, which no launcher option starts with."""
    import importlib.util
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "8", "--rows", "2000000", "--dim", "50", "--k", "15", "--self-queries"])
    a = bench.parse_args()
    assert a.n == 2_000_000 and a.gpus == 8 and a.self_queries and a.k == 15
    monkeypatch.setattr(sys, "argv", ["bench.py", "--n", "1234"])
    assert bench.parse_args().n == 1234
    from torch.distributed.run import get_args_parser
    launcher = {s for act in get_args_parser()._actions for s in act.option_strings}
    for flag in ("--rows", "--gpus", "--steps", "--warmup", "--workload", "--dtype", "--nprobe", "--self-queries", "--impl"):
        assert not any(opt.startswith(flag) for opt in launcher), flag
