"""CPU statement of the 3xFP16 operand forms of the tensor kernels: numpy float16 pieces, float64
accumulation.  It pins the REPRESENTATION part of the certificate's error model -- what the split, the scales and the unit-row
division lose before any tensor-core accumulation -- on data with a wide dynamic range:

  * queries: every row scaled by its own power of two (largest element in [2^13, 2^14)), hi = fp16(x s), lo = fp16(x s - hi);
  * L2 database operand: ONE power of two for the whole index (largest element of the index in [2^13, 2^14)); the three-term
    product s = Qhi.Xhi + Qlo.Xhi + Qhi.Xlo must match q.x to 3 * 2^-22 of |q| * |x|max (the bound eps (|q| + |x|max)^2 is stated
    in the largest row norm);
  * flat cosine operand: unit rows x / |x| (float32 division) at the uniform scale 2^13 and the negated query: the accumulator
    times the query's inverse scale / 2^13 is -q.x / |x| to (3 * 2^-22 + 2 * 2^-24) |q|.

The GPU tests assert the same model on the kernels' own tile dumps (tests/test_gpu_tensor.py::test_f32_operand_forms_agree)."""
import numpy as np
import pytest

from annb200 import datagen


def _pow2_scale(m):
    """power of two s with m * s in [2^13, 2^14) (split_f16_kernel / tc_uniform_f16_scale)"""
    m = np.asarray(m, dtype=np.float64)
    ex = np.frexp(np.where(m > 0, m, 1.0))[1]
    return np.where(m > 0, np.ldexp(1.0, np.clip(14 - ex, -100, 100)), 1.0)


def _split(xs):
    xs = xs.astype(np.float32)
    hi = xs.astype(np.float16)
    lo = (xs - hi.astype(np.float32)).astype(np.float16)
    assert np.isfinite(hi.astype(np.float32)).all()
    return hi.astype(np.float64), lo.astype(np.float64)


def _wide_range_data(seed=71, n=4000, dim=96, nq=64):
    rng = np.random.default_rng(seed)
    base = datagen.correlated(n, dim, seed=seed)
    row_scale = np.float32(10.0) ** rng.integers(-6, 7, n).astype(np.float32)             # |x| over 12 decades
    col_scale = np.float32(2.0) ** rng.integers(-10, 11, dim).astype(np.float32)          # elements over 6 decades inside a row
    data = np.ascontiguousarray(base * row_scale[:, None] * col_scale[None, :], dtype=np.float32)
    q = datagen.subsample_with_noise(data, nq, seed=seed)
    return data, q


def test_l2_uniform_scale_error_is_relative_to_the_largest_row():
    data, q = _wide_range_data()
    S = float(_pow2_scale(np.abs(data).max()))
    assert 2.0 ** 13 <= np.abs(data).max() * S < 2.0 ** 14
    xh, xl = _split(data * np.float32(S))
    sq = _pow2_scale(np.abs(q).max(axis=1))
    qh, ql = _split(q * sq[:, None].astype(np.float32))
    s = (qh + ql) @ xh.T + qh @ xl.T                         # the three product terms; lo.lo is dropped
    dot = s / (S * sq[:, None])
    x64, q64 = data.astype(np.float64), q.astype(np.float64)
    true = q64 @ x64.T
    qn, xn = np.sqrt((q64 ** 2).sum(1)), np.sqrt((x64 ** 2).sum(1))
    err = np.abs(dot - true) / (qn[:, None] * xn.max())
    assert err.max() <= 3 * 2.0 ** -22, err.max()
    # ... and NOT relative to the row's own norm: small rows are exact only through the re-rank (documented, checked on the GPU)
    own = np.abs(dot - true) / (qn[:, None] * xn[None, :])
    assert own.max() > 2.0 ** -11


@pytest.mark.parametrize("dim", [50, 96, 128])
def test_cosine_unit_rows_accumulator_is_the_selection_value(dim):
    data, q = _wide_range_data(seed=5, dim=dim)
    norms = np.sqrt((data.astype(np.float64) ** 2).sum(1)).astype(np.float32)          # the index norms (f32)
    unit = (data / norms[:, None]).astype(np.float32) * np.float32(8192.0)             # one f32 division per element, exact scale
    xh, xl = _split(unit)
    sq = _pow2_scale(np.abs(q).max(axis=1))
    qh, ql = _split(-(q * sq[:, None].astype(np.float32)))                              # negated query operand (exact)
    acc = (qh + ql) @ xh.T + qh @ xl.T
    v = acc / (sq[:, None] * 8192.0)                                                    # cq = q_inv_scale / 2^13, a power of two
    q64 = q.astype(np.float64)
    true = -(q64 @ data.astype(np.float64).T) / norms.astype(np.float64)[None, :]
    qn = np.sqrt((q64 ** 2).sum(1))
    err = np.abs(v - true) / qn[:, None]
    assert err.max() <= 3 * 2.0 ** -22 + 2 * 2.0 ** -24, err.max()


def test_padding_and_zero_norm_rows_never_win():
    """A NaN in the first element of the hi piece makes the whole accumulator column NaN; minimum / comparison ignore it."""
    acc = np.array([[-3.0, np.nan, -1.0, np.nan]])
    assert np.fmin.reduce(acc, axis=1)[0] == -3.0 and not (acc[0, 1] < -1e30)
