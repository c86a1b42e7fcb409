"""Bench setup on the GPU (not timed, not the product): synthetic data at 10M x 128 and an IVF index built with the
library's own kernels (coarse assignment = flat tensor-core search with k = 1 over the centroid table).

torch is plumbing here: random numbers, sort, gather.  The structure / constants of the generators are those of
oracle/datagen.py (examples/commons/mod.rs:339-441, 859-883); torch's Philox stream replaces numpy's PCG64, which itself
stands in for rand's StdRng.  Index contents produced here are fed identically to libannb200 and (on request) to the
CPU oracle, so parity is judged on shared contents.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)

import annb200  # noqa: E402
from annb200 import datagen  # noqa: E402


def correlated_gpu(n, dim, device, seed=42, n_clusters=datagen.DEFAULT_N_CLUSTERS, chunk=1 << 20):
    """generate_clustered_data_high_dim on the device: cluster structure drawn on the host exactly as
    oracle/datagen.correlated does (same numpy stream for centres / bases / spectra / assignments), the per-sample
    Gaussians drawn on the GPU."""
    rng = np.random.Generator(np.random.PCG64(seed))
    scale = np.sqrt(dim) * 2.0
    min_sep = scale * 0.8
    centres = []
    while len(centres) < n_clusters:
        cand = rng.uniform(-scale, scale, dim)
        if all(((cand - c) ** 2).sum() >= min_sep ** 2 for c in centres):
            centres.append(cand)
    corr_rank = min(datagen.DEFAULT_CORR_RANK, dim)
    gbasis = datagen._orthonormal_basis(rng, dim, corr_rank)
    gspec = (scale / 10.0) / (np.arange(1, corr_rank + 1) ** datagen.DEFAULT_ANISO_DECAY)
    rank = min(datagen.DEFAULT_LOCAL_RANK, dim)
    bases = np.asarray([datagen._orthonormal_basis(rng, dim, rank) for _ in range(n_clusters)])
    spectra = np.asarray([rng.uniform(0.3, 1.0) * scale / 10.0 / (np.arange(1, rank + 1) ** datagen.DEFAULT_ANISO_DECAY)
                          for _ in range(n_clusters)])
    floor = scale / 100.0
    cs = datagen.DEFAULT_COR_STRENGTH
    sg, sl = np.sqrt(cs), np.sqrt(1.0 - cs)
    assign = datagen._cluster_assignments(rng, n, n_clusters)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    t_cent = torch.tensor(np.asarray(centres), dtype=torch.float32, device=device)
    t_gb = torch.tensor(gbasis, dtype=torch.float32, device=device)
    t_gs = torch.tensor(gspec * sg, dtype=torch.float32, device=device)
    t_bases = torch.tensor(bases, dtype=torch.float32, device=device)          # [C, dim, rank]
    t_spec = torch.tensor(spectra * sl, dtype=torch.float32, device=device)    # [C, rank]
    t_assign = torch.from_numpy(assign).to(device)
    out = torch.empty((n, dim), dtype=torch.float32, device=device)
    for s in range(0, n, chunk):
        a = t_assign[s:s + chunk]
        m = a.numel()
        x = t_cent[a] + torch.randn((m, dim), generator=g, device=device) * floor
        x += (torch.randn((m, corr_rank), generator=g, device=device) * t_gs) @ t_gb.T
        zl = torch.randn((m, rank), generator=g, device=device) * t_spec[a]
        for c in range(n_clusters):
            sel = (a == c).nonzero(as_tuple=True)[0]
            if sel.numel():
                x[sel] += zl[sel] @ t_bases[c].T
        out[s:s + m] = x
    return out


def subsample_with_noise_gpu(data, nq, seed=42):
    g = torch.Generator(device=data.device)
    g.manual_seed(seed + 1000)
    idx = torch.randperm(data.shape[0], generator=g, device=data.device)[:nq]
    return (data[idx] + torch.randn((nq, data.shape[1]), generator=g, device=data.device) * 0.05).contiguous()


def _flat_handle_from_device(t, metric, dtype, device_index, id_base=0):
    torch.cuda.synchronize(t.device)      # the library copies on its own stream: the tensor must be complete
    h = C.c_void_p()
    annb200._check(annb200.lib().annb_flat_create(C.byref(h), C.c_void_p(t.data_ptr()), t.shape[0], t.shape[1], dtype, metric, None, id_base, device_index))
    return annb200.ExhaustiveIndexB200(h)


def nearest_centroid(points, centroids, device_index, chunk=1 << 20):
    """argmin_c |x - c|^2 for every row of `points` with the library's flat search (k = 1) over the centroid table."""
    lib = annb200.lib()
    cix = _flat_handle_from_device(centroids.contiguous(), annb200.L2, annb200.F32, device_index)
    n, dim = points.shape
    out = torch.empty((n,), dtype=torch.int64, device=points.device)
    ids = torch.empty((min(chunk, n), 1), dtype=torch.int64, device=points.device)
    st = torch.cuda.current_stream(points.device).cuda_stream
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        annb200._check(lib.annb_flat_search_dev(cix.handle, points[s:s + m].data_ptr(), m, dim, 1, ids.data_ptr(), None, None, st))
        out[s:s + m] = ids[:m, 0]
    torch.cuda.synchronize(points.device)
    cix.close()
    return out


def train_centroids_gpu(train, nlist, iters, device_index):
    """Plain Lloyd (stand-in for train_centroids, k_means_utils.rs:2771-2938; empty clusters keep their centroid)."""
    n = train.shape[0]
    cent = train[(torch.arange(nlist, device=train.device) * n) // nlist].clone()
    for _ in range(iters):
        a = nearest_centroid(train, cent, device_index)
        sums = torch.zeros_like(cent, dtype=torch.float64)
        sums.index_add_(0, a, train.double())
        cnt = torch.bincount(a, minlength=nlist)
        nz = cnt > 0
        cent[nz] = (sums[nz] / cnt[nz].unsqueeze(1)).float()
    return cent


def build_ivf_parts_gpu(data, nlist, dtype, device_index, seed=42, kmeans_iters=8):
    """Steps of IvfIndex::build (src/cpu/ivf.rs:145-249) for the L2 metric, on the device.  Returns the contents of the
    reference's index struct: list-ordered vectors (tensor, index dtype), centroids, CSR offsets, original ids."""
    n, dim = data.shape
    g = torch.Generator(device=data.device)
    g.manual_seed(seed)
    n_train = max(min(256 * nlist, 250_000, n), 1)                      # ivf.rs:174
    train = data[torch.randperm(n, generator=g, device=data.device)[:n_train]].contiguous()
    cent = train_centroids_gpu(train, nlist, kmeans_iters, device_index)
    assign = nearest_centroid(data, cent, device_index)
    order = torch.sort(assign, stable=True).indices                    # build_csr_layout: stable counting sort
    counts = torch.bincount(assign, minlength=nlist).cpu().numpy()
    offsets = np.zeros(nlist + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(counts)
    base = dict(vectors=data[order].contiguous(), centroids=cent, offsets=offsets, original_ids=order.contiguous(), scales=None,
                train_absmax=train.abs().amax(dim=0))
    return requantise_parts(base, dtype)


def requantise_parts(base, dtype, chunk=1 << 20):
    """The BF16 / SQ8 form of an f32 parts dict (same lists, same centroids): encode_bf16_quantisation (quantisers.rs:31-38)
    or ScalarQuantiser with the codebook of the training sample (ivf_sq8.rs:211).  F32: the dict itself."""
    if dtype == annb200.F32:
        return base
    src = base["vectors"]
    out = dict(base)
    if dtype == annb200.BF16:
        out["vectors"] = src.to(torch.bfloat16).contiguous()              # RNE, as half::bf16::from_f32
        return out
    mx = base["train_absmax"]
    scales = torch.where(mx <= 0, torch.ones_like(mx), mx / 128.0)
    vec = torch.empty(src.shape, dtype=torch.int8, device=src.device)
    for s in range(0, src.shape[0], chunk):
        scaled = src[s:s + chunk] / scales
        r = scaled + 0.5 * torch.where(torch.signbit(scaled), -torch.ones_like(scaled), torch.ones_like(scaled))
        vec[s:s + chunk] = torch.trunc(r.clamp(-128.0, 127.0)).to(torch.int8)
    out["vectors"] = vec
    out["scales"] = scales
    return out


def ivf_handle_from_parts(parts, n_total, dim, dtype, metric, device_index, list_begin=0, list_end=None):
    """device_index: one ordinal (optionally with a list range = one shard), or a list of ordinals (annb_ivf_create_multi)."""
    lib = annb200.lib()
    nlist = parts["centroids"].shape[0]
    off = parts["offsets"]
    if isinstance(device_index, (list, tuple)):
        vec, oid, cent, sc = parts["vectors"], parts["original_ids"], parts["centroids"].contiguous(), parts["scales"]
        torch.cuda.synchronize(vec.device)
        devs = (C.c_int * len(device_index))(*[int(d) for d in device_index])
        h = C.c_void_p()
        annb200._check(lib.annb_ivf_create_multi(C.byref(h), C.c_void_p(vec.data_ptr()), None, C.c_void_p(cent.data_ptr()), None,
                                                 C.c_void_p(off.ctypes.data), C.c_void_p(oid.data_ptr()), n_total, dim, nlist, dtype, metric,
                                                 None if sc is None else C.c_void_p(sc.contiguous().data_ptr()), devs, len(device_index)))
        return annb200.IvfIndexB200(h)
    list_end = nlist if list_end is None else list_end
    r0, r1 = int(off[list_begin]), int(off[list_end])
    vec = parts["vectors"][r0:r1]
    oid = parts["original_ids"][r0:r1]
    cent = parts["centroids"].contiguous()
    sc = parts["scales"]
    torch.cuda.synchronize(vec.device)    # the library copies on its own stream: the tensors must be complete
    h = C.c_void_p()
    annb200._check(lib.annb_ivf_create(C.byref(h), C.c_void_p(vec.data_ptr()), None, C.c_void_p(cent.data_ptr()), None,
                                       C.c_void_p(off.ctypes.data), C.c_void_p(oid.data_ptr()), n_total, dim, nlist, dtype, metric,
                                       None if sc is None else C.c_void_p(sc.contiguous().data_ptr()), list_begin, list_end, device_index))
    return annb200.IvfIndexB200(h)


def exact_ground_truth(data, queries, k, metric, device_index):
    """Exact f32 top-k with the library's flat index (tensor path + exact re-rank)."""
    lib = annb200.lib()
    ix = _flat_handle_from_device(data, metric, annb200.F32, device_index)
    nq = queries.shape[0]
    ids = torch.empty((nq, k), dtype=torch.int64, device=data.device)
    st = torch.cuda.current_stream(data.device).cuda_stream
    annb200._check(lib.annb_flat_search_dev(ix.handle, queries.data_ptr(), nq, queries.shape[1], k, ids.data_ptr(), None, None, st))
    torch.cuda.synchronize(data.device)
    ix.close()
    return ids.cpu().numpy()
