"""Neighbour-gap statistics of the flat bench workload: how far the 65th / 150th neighbour lies beyond the 10th, in the
units the coverage certificate tests (squared distance for L2, cosine distance) -- i.e. how large an error bound the
second-chance certificate can absorb."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python"), os.path.join(ROOT, "tools")]
import annb200
import gpu_setup as gs
dev = torch.device("cuda:0")
n, dim, nq, k = 1_000_000, 128, 2000, 160
data = gs.correlated_gpu(n, dim, dev, seed=42)
q = gs.subsample_with_noise_gpu(data, 10_000, seed=42)[:nq].contiguous()
lib = annb200.lib()
st = torch.cuda.current_stream(dev).cuda_stream
xn = data.norm(dim=1)
print(f"row norms: min {xn.min().item():.1f} median {xn.median().item():.1f} max {xn.max().item():.1f}; query norm median {q.norm(dim=1).median().item():.1f}")
for metric, name in ((annb200.L2, "l2"), (annb200.COSINE, "cosine")):
    ix = gs._flat_handle_from_device(data, metric, annb200.BF16, 0)
    ix.set_option("path", annb200.PATH_SIMT)
    d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    annb200._check(lib.annb_flat_search_dev(ix.handle, q.data_ptr(), nq, dim, k, ids.data_ptr(), d.data_ptr(), None, st))
    torch.cuda.synchronize()
    d = d.cpu().numpy().astype(np.float64)
    qn = q.norm(dim=1).cpu().numpy().astype(np.float64)
    for eps in (2.9e-6, 1.0e-5, 1.9e-5):
        margin = eps * (qn + float(xn.max())) ** 2 if name == "l2" else np.full(nq, eps)
        for j in (16, 32, 64, 150):
            frac = float(((d[:, j] - d[:, 9]) <= margin).mean())
            print(f"{name} eps {eps:.1e}: gap(d[{j}] - d[9]) <= margin for {frac:.4f} of the queries (median gap {np.median(d[:, j] - d[:, 9]):.3e}, median margin {np.median(margin):.3e}, median d[9] {np.median(d[:, 9]):.3e})")
    ix.close()
