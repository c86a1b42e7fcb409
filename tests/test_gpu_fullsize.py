"""Full-size, size-independent properties (BASELINE.json configs 2, 3 and 5) on data generated on the device:
the oracle cannot finish these sizes in seconds, so the checks are (i) the two independent GPU paths agree bit for
bit, (ii) sortedness / self-at-rank-0 / idempotence, (iii) a small oracle sample."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import annb200
from oracle import oracle as o
from util import assert_exact

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def _search_dev(ix, q_t, k, ivf_nprobe=None):
    import torch
    lib = annb200.lib()
    nq, dim = q_t.shape
    ids = torch.empty((nq, k), dtype=torch.int64, device=q_t.device)
    d = torch.empty((nq, k), dtype=torch.float32, device=q_t.device)
    st = torch.cuda.current_stream().cuda_stream
    if ivf_nprobe is None:
        annb200._check(lib.annb_flat_search_dev(ix.handle, q_t.data_ptr(), nq, dim, k, ids.data_ptr(), d.data_ptr(), None, st))
    else:
        annb200._check(lib.annb_ivf_search_dev(ix.handle, q_t.data_ptr(), nq, dim, k, ivf_nprobe, ids.data_ptr(), d.data_ptr(), None, st))
    torch.cuda.synchronize()
    return ids.cpu().numpy(), d.cpu().numpy()


def test_config2_flat_1m_x_128_cosine_tensor_equals_exact_path(gpu):
    import torch
    import gpu_setup as gs
    dev = torch.device("cuda:0")
    data = gs.correlated_gpu(1_000_000, 128, dev, seed=42)
    q = gs.subsample_with_noise_gpu(data, 1024, seed=42)
    ix = gs._flat_handle_from_device(data, annb200.COSINE, annb200.F32, 0)
    ix.set_option("path", annb200.PATH_TENSOR)
    a = _search_dev(ix, q, 10)
    b = _search_dev(ix, q, 10)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), "not idempotent"
    ix.set_option("path", annb200.PATH_SIMT)
    c = _search_dev(ix, q, 10)
    assert_exact(a[0], a[1], c[0], c[1], "1M x 128 cosine: tensor path vs exact CUDA-core path")
    assert (np.diff(a[1], axis=1) >= 0).all()
    # oracle on a few queries (host copy of the data)
    host = data.cpu().numpy()
    ref = o.flat_search(o.build_flat(host, o.COSINE), q[:16].cpu().numpy(), 10)
    assert_exact(a[0][:16], a[1][:16], ref[0], ref[1], "1M x 128 cosine vs oracle sample")


def test_config5_self_knn_dim50_k15(gpu):
    """Self-kNN at d = 50 (padded to 64 on the tensor path), k = 15 -> k' = 32; a 200k slice of config 5."""
    import torch
    import gpu_setup as gs
    dev = torch.device("cuda:0")
    data = gs.correlated_gpu(200_000, 50, dev, seed=7)
    ix = gs._flat_handle_from_device(data, annb200.L2, annb200.F32, 0)
    n, k = 4096, 15
    ids = np.empty((n, k), np.uint64); d = np.empty((n, k), np.float32); cnt = np.empty(n, np.uint32)
    lib = annb200.lib()
    ix.set_option("path", annb200.PATH_TENSOR)
    annb200._check(lib.annb_flat_search_self(ix.handle, 1000, 1000 + n, k, C.c_void_p(ids.ctypes.data), C.c_void_p(d.ctypes.data), C.c_void_p(cnt.ctypes.data)))
    assert ix.get_stat("last_path") == annb200.PATH_TENSOR
    assert (ids[:, 0] == np.arange(1000, 1000 + n)).all() and (d[:, 0] == 0).all(), "self must be at rank 0 with distance 0"
    ids2 = np.empty_like(ids); d2 = np.empty_like(d)
    ix.set_option("path", annb200.PATH_SIMT)
    annb200._check(lib.annb_flat_search_self(ix.handle, 1000, 1000 + n, k, C.c_void_p(ids2.ctypes.data), C.c_void_p(d2.ctypes.data), C.c_void_p(cnt.ctypes.data)))
    assert_exact(ids.view(np.int64), d, ids2.view(np.int64), d2, "self-kNN d=50: tensor vs exact path")


@pytest.mark.parametrize("dtype", ["f32", "bf16", "sq8"])
def test_config3_4_ivf_2m_x_128_scan_variants_agree(gpu, dtype):
    """IVF at 2M x 128, nlist 1024: the list-major and the query-major scan kernels are independent implementations
    of the same (distance, position) selection -- their outputs must be identical; recall against the exact flat
    search must be in the expected range."""
    import torch
    import gpu_setup as gs
    dev = torch.device("cuda:0")
    dt = {"f32": annb200.F32, "bf16": annb200.BF16, "sq8": annb200.SQ8}[dtype]
    data = gs.correlated_gpu(2_000_000, 128, dev, seed=11)
    q = gs.subsample_with_noise_gpu(data, 512, seed=11)
    parts = gs.build_ivf_parts_gpu(data, 1024, dt, 0, seed=11, kmeans_iters=4)
    ix = gs.ivf_handle_from_parts(parts, data.shape[0], 128, dt, annb200.L2, 0)
    ix.set_option("ivf_list_major", 1)
    a = _search_dev(ix, q, 10, ivf_nprobe=16)
    ix.set_option("ivf_list_major", 0)
    b = _search_dev(ix, q, 10, ivf_nprobe=16)
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    assert np.array_equal(a[0], b[0])
    assert (np.diff(a[1], axis=1) >= 0).all() and (a[0] >= 0).all()
    assert ix.get_stat("coarse_path") == 2                      # both runs ranked the centroids on the tensor cores
    # ... and the whole pipeline on the exact CUDA-core kernels (exact centroid ranking, list-major CUDA-core scan)
    ix.set_option("ivf_tc_coarse", 0)
    ix.set_option("ivf_list_major", 1)
    ix.set_option("path", annb200.PATH_SIMT)
    c = _search_dev(ix, q, 10, ivf_nprobe=16)
    assert ix.get_stat("coarse_path") != 2 and ix.get_stat("last_path") == annb200.PATH_SIMT
    assert np.array_equal(a[1].view(np.uint32), c[1].view(np.uint32))
    assert np.array_equal(a[0], c[0])
    truth = gs.exact_ground_truth(data, q, 10, annb200.L2, 0)
    rec = o.recall_at_k(truth, a[0], 10)
    assert rec > {"f32": 0.9, "bf16": 0.85, "sq8": 0.5}[dtype], rec


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[2] / [3] at full size against the CPU oracle: IVF 10M x 128, nlist 4096, nprobe 32, 10k-query batch.
# The index is built once on the device (tools/gpu_setup.py); its contents are copied to the host once per dtype and fed
# to the oracle unchanged (parity on shared index contents).
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ivf_10m(gpu):
    import torch
    import gpu_setup as gs
    dev = torch.device("cuda:0")
    data = gs.correlated_gpu(10_000_000, 128, dev, seed=42)
    q = gs.subsample_with_noise_gpu(data, 10_000, seed=42)
    base = gs.build_ivf_parts_gpu(data, 4096, annb200.F32, 0, seed=42, kmeans_iters=4)
    del data
    torch.cuda.empty_cache()
    yield {"base": base, "q": q}
    del base
    torch.cuda.empty_cache()


@pytest.mark.parametrize("dtype", ["f32", "bf16", "sq8"])
def test_config3_4_ivf_10m_x_128_matches_oracle(gpu, ivf_10m, dtype):
    import torch
    import gpu_setup as gs
    dt = {"f32": annb200.F32, "bf16": annb200.BF16, "sq8": annb200.SQ8}[dtype]
    odt = {"f32": o.F32, "bf16": o.BF16, "sq8": o.SQ8}[dtype]
    n, dim, nlist, nprobe, k = 10_000_000, 128, 4096, 32, 10
    parts = gs.requantise_parts(ivf_10m["base"], dt)
    ix = gs.ivf_handle_from_parts(parts, n, dim, dt, annb200.L2, 0)
    q = ivf_10m["q"]
    ids, d = _search_dev(ix, q, k, ivf_nprobe=nprobe)                       # the whole 10k batch, as the bench runs it
    assert ix.get_stat("last_path") == annb200.PATH_TENSOR and ix.get_stat("coarse_path") == 2
    vec = parts["vectors"]
    vec_np = vec.view(torch.int16).cpu().numpy().view(np.uint16) if dt == annb200.BF16 else vec.cpu().numpy()
    oi = o.IvfIndex(odt, o.L2, n, dim, nlist, vec_np, parts["centroids"].cpu().numpy(), parts["offsets"].astype(np.int64),
                    parts["original_ids"].cpu().numpy(), scales=None if parts["scales"] is None else parts["scales"].cpu().numpy())
    ns = 96
    ref = o.ivf_search(oi, q[:ns].cpu().numpy(), k, nprobe=nprobe)
    if dtype == "sq8":
        from util import assert_tie_classes
        assert_tie_classes(ids[:ns], d[:ns], ref[0], ref[1], "IVF-SQ8 10M vs oracle")
    else:
        assert_exact(ids[:ns], d[:ns], ref[0], ref[1], f"IVF-{dtype} 10M x 128 nprobe 32 vs oracle")
    # the same handle through the host-buffer entry point and split over four shards on this device (one process)
    h_ids, h_d, _ = ix.query_batch(q[:2048].cpu().numpy(), k, nprobe=nprobe)
    assert np.array_equal(h_ids, ids[:2048]) and np.array_equal(h_d.view(np.uint32), d[:2048].view(np.uint32))
    ix.close()
    if dtype == "f32":
        m = gs.ivf_handle_from_parts(parts, n, dim, dt, annb200.L2, [0, 0, 0, 0])
        m_ids, m_d, _ = m.query_batch(q[:2048].cpu().numpy(), k, nprobe=nprobe)
        assert np.array_equal(m_ids, ids[:2048]) and np.array_equal(m_d.view(np.uint32), d[:2048].view(np.uint32)), "4 list shards vs unsharded"
        m.close()


def test_config5_2m_x_50_k15_eight_shards_vs_unsharded_and_oracle(gpu):
    """BASELINE configs[4]: self-kNN 2M x 50, k = 15, database rows split into 8 shards (here: 8 handles on one device, the
    exchange step through annb_merge_shards_dev exactly as the 8-rank job runs it), against the unsharded handle, the
    multi-device handle and an oracle sample."""
    import torch
    import gpu_setup as gs
    dev = torch.device("cuda:0")
    n, dim, k, nq, parts = 2_000_000, 50, 15, 4096, 8
    data = gs.correlated_gpu(n, dim, dev, seed=42)
    q = data[:nq].contiguous()
    lib = annb200.lib()
    st = torch.cuda.current_stream().cuda_stream
    full = gs._flat_handle_from_device(data, annb200.L2, annb200.F32, 0)
    f_ids, f_d = _search_dev(full, q, k)
    assert (f_ids[:, 0] == np.arange(nq)).all() and (f_d[:, 0] == 0).all()
    block = ((nq * k * 12 + 255) // 256) * 256
    gathered = torch.empty((parts * block,), dtype=torch.uint8, device=dev)
    for p in range(parts):
        lo, hi = (p * n) // parts, ((p + 1) * n) // parts
        sh = gs._flat_handle_from_device(data[lo:hi], annb200.L2, annb200.F32, 0, id_base=lo)
        slot = gathered[p * block:(p + 1) * block]
        annb200._check(lib.annb_flat_search_dev(sh.handle, q.data_ptr(), nq, dim, k, slot.data_ptr(), slot[nq * k * 8:].data_ptr(), None, st))
        torch.cuda.synchronize()
        sh.close()
    m_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    m_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    annb200._check(lib.annb_merge_shards_dev(gathered.data_ptr(), block, nq * k * 8, parts, nq, k, m_ids.data_ptr(), m_d.data_ptr(), None, st))
    torch.cuda.synchronize()
    assert_exact(m_ids.cpu().numpy(), m_d.cpu().numpy(), f_ids, f_d, "8 row shards merged vs unsharded")
    full.close()
    host = data.cpu().numpy()
    del data, gathered
    torch.cuda.empty_cache()
    ref = o.flat_search(o.build_flat(host, o.L2), None, k, self_rows=np.arange(64), self_mode=True)
    assert_exact(f_ids[:64], f_d[:64], ref[0], ref[1], "2M x 50 self-kNN vs oracle sample")
    multi = annb200.ExhaustiveIndexB200.new(host, annb200.L2, annb200.F32, device=[i % gpu for i in range(8)])
    g_ids, g_d, _ = multi.generate_knn(k, row_begin=0, row_end=nq)
    assert_exact(g_ids, g_d, f_ids, f_d, "multi-device handle (8 shards) generate_knn vs unsharded")
    multi.close()
