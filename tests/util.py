"""Comparison helpers shared by the parity tests."""
import numpy as np


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_exact(ids, dist, ref_ids, ref_dist, what=""):
    """Bit-exact distances and identical ids.  Both sides order by (distance, id), the total order of the
    reference's (OrderedFloat, usize) tuples, so no tolerance is needed."""
    ids = np.asarray(ids)
    ref_ids = np.asarray(ref_ids)
    assert ids.shape == ref_ids.shape, what
    bad = np.nonzero((ids != ref_ids).any(axis=1) | (bits(dist) != bits(ref_dist)).any(axis=1))[0]
    if bad.size:
        r = bad[0]
        raise AssertionError(f"{what}: {bad.size}/{ids.shape[0]} rows differ; first row {r}\n gpu ids {ids[r]}\n ref ids {ref_ids[r]}\n"
                             f" gpu d {np.asarray(dist)[r]}\n ref d {np.asarray(ref_dist)[r]}")


def assert_tie_classes(ids, dist, ref_ids, ref_dist, what=""):
    """Distances bit-exact; ids equal as sets inside every run of equal distance that ends before the
    last valid slot.  The last run may be cut by k, where the reference's heap keeps an
    implementation-defined subset (src/quantised/ivf_sq8.rs:329-352): there only check that the ids are
    distinct and valid."""
    ids = np.asarray(ids)
    ref_ids = np.asarray(ref_ids)
    assert (bits(dist) == bits(ref_dist)).all(), f"{what}: distances differ"
    for r in range(ids.shape[0]):
        d = np.asarray(dist)[r]
        valid = int((ids[r] >= 0).sum())
        assert valid == int((ref_ids[r] >= 0).sum()), f"{what}: row {r} count"
        j = 0
        while j < valid:
            e = j
            while e + 1 < valid and d[e + 1] == d[j]:
                e += 1
            if e < valid - 1:
                assert set(ids[r, j:e + 1].tolist()) == set(ref_ids[r, j:e + 1].tolist()), f"{what}: row {r} run [{j},{e}]"
            else:
                assert len(set(ids[r, j:e + 1].tolist())) == e + 1 - j, f"{what}: row {r} duplicate ids in last run"
            j = e + 1


def assert_tolerance(ids, dist, ref_ids, ref_dist, rtol=1e-5, atol=1e-5, what=""):
    """The north-star contract: index sets agree except for swaps among candidates whose distances differ
    by less than the tolerance (1e-5 relative, with an absolute floor for distances near zero), and
    distances agree within that tolerance."""
    ids = np.asarray(ids)
    ref_ids = np.asarray(ref_ids)
    dist = np.asarray(dist, dtype=np.float64)
    ref_dist = np.asarray(ref_dist, dtype=np.float64)
    fin = np.isfinite(ref_dist)
    assert (np.isfinite(dist) == fin).all(), f"{what}: validity pattern differs"
    tol = atol + rtol * np.abs(ref_dist[fin])
    assert (np.abs(dist[fin] - ref_dist[fin]) <= tol).all(), f"{what}: distance tolerance exceeded"
    for r in range(ids.shape[0]):
        a, b = set(ids[r][fin[r]].tolist()), set(ref_ids[r][fin[r]].tolist())
        if a == b:
            continue
        # ids only in one list must sit within tolerance of the k-th distance
        kth = ref_dist[r][fin[r]].max()
        lim = atol + rtol * abs(kth)
        for i, v in enumerate(ids[r]):
            if fin[r, i] and v not in b:
                assert abs(dist[r, i] - kth) <= lim, f"{what}: row {r} id {v} is not a near-tie of the k-th neighbour"
        for i, v in enumerate(ref_ids[r]):
            if fin[r, i] and v not in a:
                assert abs(ref_dist[r, i] - kth) <= lim, f"{what}: row {r} ref id {v} is not a near-tie of the k-th neighbour"
