"""Synthetic inputs: numpy restatement of examples/commons/mod.rs generators.

Inputs only: used by bench.py, tools/ and tests/ (no search arithmetic lives here).  Structure and constants follow
/root/reference/examples/commons/mod.rs; the random stream is numpy's PCG64
because rand 0.9's StdRng (ChaCha12) stream cannot be reproduced without a Rust
toolchain -- the *distribution* matches, individual samples do not.

  gaussian_noise  -> generate_clustered_data           (mod.rs:174-269)
  correlated      -> generate_clustered_data_high_dim  (mod.rs:339-441)
  subsample_with_noise -> subsample_with_noise         (mod.rs:859-883)
"""
from __future__ import annotations

import numpy as np

DEFAULT_N_CLUSTERS = 25          # mod.rs:23
DEFAULT_SEED = 42                # mod.rs:27
DEFAULT_COR_STRENGTH = 0.5       # mod.rs:31
DEFAULT_BRIDGE_FRACTION = 0.2    # mod.rs:37
DEFAULT_LOCAL_RANK = 16          # mod.rs:41
DEFAULT_CORR_RANK = 32           # mod.rs:43
DEFAULT_ANISO_DECAY = 1.0        # mod.rs:45


def _cluster_assignments(rng, n_budget, n_clusters):
    assign = []
    for c in range(n_clusters):
        w = rng.uniform(0.5, 2.5)
        assign.append(np.full(int((n_budget * w) / (n_clusters * 1.25)), c, dtype=np.int64))
    assign = np.concatenate(assign) if assign else np.zeros(0, dtype=np.int64)
    if assign.size < n_budget:
        assign = np.concatenate([assign, rng.integers(0, n_clusters, n_budget - assign.size)])
    rng.shuffle(assign)
    return assign[:n_budget]


def gaussian_noise(n_samples, dim, n_clusters=DEFAULT_N_CLUSTERS, seed=DEFAULT_SEED, chunk=1 << 18):
    rng = np.random.Generator(np.random.PCG64(seed))
    centres = np.empty((n_clusters, dim))
    stds = np.empty(n_clusters)
    for c in range(n_clusters):
        centres[c] = rng.uniform(-7.5, 7.5, dim)
        stds[c] = rng.uniform(0.5, 2.5)
    edges = []
    if n_clusters >= 2:
        d2 = ((centres[:, None, :] - centres[None, :, :]) ** 2).sum(-1)
        np.fill_diagonal(d2, np.inf)
        for a in range(n_clusters):
            b = int(np.argmin(d2[a]))
            e = (min(a, b), max(a, b))
            if e not in edges:
                edges.append(e)
    n_bridge = int(n_samples * DEFAULT_BRIDGE_FRACTION) if edges else 0
    n_blob = n_samples - n_bridge
    assign = _cluster_assignments(rng, n_blob, n_clusters)
    data = np.empty((n_samples, dim), dtype=np.float32)
    for s in range(0, n_blob, chunk):
        a = assign[s:s + chunk]
        data[s:s + a.size] = (centres[a] + rng.standard_normal((a.size, dim)) * stds[a][:, None]).astype(np.float32)
    if n_bridge:
        e = np.asarray(edges)[rng.integers(0, len(edges), n_bridge)]
        t = rng.random(n_bridge)
        tube = (stds[e[:, 0]] + stds[e[:, 1]]) * 0.5 * 0.3
        for s in range(0, n_bridge, chunk):
            sl = slice(s, min(s + chunk, n_bridge))
            mid = (1.0 - t[sl])[:, None] * centres[e[sl, 0]] + t[sl][:, None] * centres[e[sl, 1]]
            m = mid.shape[0]
            data[n_blob + s:n_blob + s + m] = (mid + rng.standard_normal((m, dim)) * tube[sl][:, None]).astype(np.float32)
    return data


def _orthonormal_basis(rng, dim, rank):
    r = min(rank, dim)
    b = rng.standard_normal((dim, r)).astype(np.float32)
    for col in range(r):            # modified Gram-Schmidt as in mod.rs:283-307
        for prev in range(col):
            b[:, col] -= np.float32(np.dot(b[:, col], b[:, prev])) * b[:, prev]
        nrm = np.sqrt(np.float32(np.dot(b[:, col], b[:, col])))
        if nrm > np.finfo(np.float32).eps:
            b[:, col] /= nrm
    return b


def correlated(n_samples, dim, n_clusters=DEFAULT_N_CLUSTERS, correlation_strength=DEFAULT_COR_STRENGTH,
               seed=DEFAULT_SEED, chunk=1 << 18):
    rng = np.random.Generator(np.random.PCG64(seed))
    scale = np.sqrt(dim) * 2.0
    min_sep = scale * 0.8
    centres = []
    while len(centres) < n_clusters:
        cand = rng.uniform(-scale, scale, dim)
        if all(((cand - c) ** 2).sum() >= min_sep ** 2 for c in centres):
            centres.append(cand)
    centres = np.asarray(centres)
    corr_rank = min(DEFAULT_CORR_RANK, dim)
    gbasis = _orthonormal_basis(rng, dim, corr_rank)
    gspec = (scale / 10.0) / (np.arange(1, corr_rank + 1) ** DEFAULT_ANISO_DECAY)
    rank = min(DEFAULT_LOCAL_RANK, dim)
    bases = [_orthonormal_basis(rng, dim, rank) for _ in range(n_clusters)]
    spectra = []
    for _ in range(n_clusters):
        s = rng.uniform(0.3, 1.0) * scale / 10.0
        spectra.append(s / (np.arange(1, rank + 1) ** DEFAULT_ANISO_DECAY))
    spectra = np.asarray(spectra)
    bases = np.asarray(bases)                       # [C, dim, rank]
    floor = scale / 100.0
    cs = min(max(correlation_strength, 0.0), 1.0)
    sg, sl = np.sqrt(cs), np.sqrt(1.0 - cs)
    assign = _cluster_assignments(rng, n_samples, n_clusters)
    data = np.empty((n_samples, dim), dtype=np.float32)
    for s in range(0, n_samples, chunk):
        a = assign[s:s + chunk]
        m = a.size
        x = (centres[a] + rng.standard_normal((m, dim)) * floor).astype(np.float32)
        zg = (rng.standard_normal((m, corr_rank)) * (gspec * sg)).astype(np.float32)
        x += zg @ gbasis.T
        zl = (rng.standard_normal((m, rank)) * (spectra[a] * sl)).astype(np.float32)
        for c in np.unique(a):
            sel = np.nonzero(a == c)[0]
            x[sel] += zl[sel] @ bases[c].T
        data[s:s + m] = x
    return data


def subsample_with_noise(data, n_samples, seed=DEFAULT_SEED):
    """Queries: random row subset + N(0, 0.05^2) per coordinate (seed + 1000)."""
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    n = data.shape[0]
    m = min(n_samples, n)
    idx = rng.permutation(n)[:m]
    u1 = 1.0 - rng.random((m, data.shape[1]))      # (0, 1]: avoids ln(0)
    u2 = rng.random((m, data.shape[1]))
    noise = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    return (data[idx].astype(np.float64) + noise * 0.05).astype(np.float32)


def make(kind, n, dim, seed=DEFAULT_SEED):
    if kind in ("gaussian", "gaussian_noise"):
        return gaussian_noise(n, dim, seed=seed)
    if kind == "correlated":
        return correlated(n, dim, seed=seed)
    raise ValueError(kind)
