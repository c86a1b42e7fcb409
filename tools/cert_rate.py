"""Reports how many queries of the bench workloads pass the coverage certificate of the tensor paths."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python"), os.path.join(ROOT, "tools")]
import torch
import annb200
import gpu_setup as gs
dev = torch.device("cuda:0")
n, dim, nq = 1_000_000, 128, 10_000
data = gs.correlated_gpu(n, dim, dev, seed=42)
q = gs.subsample_with_noise_gpu(data, nq, seed=42)
lib = annb200.lib()
ids = torch.empty((nq, 10), dtype=torch.int64, device=dev); d = torch.empty((nq, 10), dtype=torch.float32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for metric, mname in ((annb200.COSINE, "cosine"), (annb200.L2, "l2")):
    for dt, dname in ((annb200.F32, "f32"), (annb200.BF16, "bf16")):
        ix = gs._flat_handle_from_device(data, metric, dt, 0)
        ix.set_option("path", annb200.PATH_TENSOR)
        for e in (-18, -20, -22):
            ix.set_option("cert_eps_log2", e)
            annb200._check(lib.annb_flat_search_dev(ix.handle, q.data_ptr(), nq, dim, 10, ids.data_ptr(), d.data_ptr(), None, st))
            torch.cuda.synchronize()
            print(f"flat {dname} {mname}: eps=2^{e}: uncertified {ix.get_stat('uncertified')} / {nq}")
        ix.close()

# IVF list scan (10M x 128 L2 unless CERT_N is set)
n2 = int(os.environ.get("CERT_N", 10_000_000))
data = None
torch.cuda.empty_cache()
data = gs.correlated_gpu(n2, dim, dev, seed=42)
q = gs.subsample_with_noise_gpu(data, nq, seed=42)
for dt, dname in ((annb200.F32, "f32"), (annb200.BF16, "bf16"), (annb200.SQ8, "sq8")):
    parts = gs.build_ivf_parts_gpu(data, 4096, dt, 0, seed=42, kmeans_iters=8)
    ix = gs.ivf_handle_from_parts(parts, n2, dim, dt, annb200.L2, 0)
    ix.set_option("cert_fallback", 0)
    for e in (-18, -20, -22):
        ix.set_option("cert_eps_log2", e)
        annb200._check(lib.annb_ivf_search_dev(ix.handle, q.data_ptr(), nq, dim, 10, 32, ids.data_ptr(), d.data_ptr(), None, st))
        torch.cuda.synchronize()
        print(f"ivf {dname} l2 nprobe=32: eps=2^{e}: uncertified {ix.get_stat('uncertified')} / {nq} (coarse path {ix.get_stat('coarse_path')})")
    ix.close()
