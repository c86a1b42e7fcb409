"""A/B of one integer index option on the flat bench workload (1M x 128 Correlated, 10k-query batch, k = 10), one index per dtype.
usage: python tools/flat_ab.py [dtypes: f32,bf16,sq8] [metric: cosine|l2] [option name] [values: 0,1,...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python"), os.path.join(ROOT, "tools")]
import annb200
import gpu_setup as gs
dts = sys.argv[1].split(",") if len(sys.argv) > 1 else ["bf16"]
metric = annb200.COSINE if (sys.argv[2] if len(sys.argv) > 2 else "cosine") == "cosine" else annb200.L2
opt = sys.argv[3] if len(sys.argv) > 3 else "tc_bf16_terms"
values = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [3, 2]
n, dim, nq, k = int(os.environ.get("AB_N", 1_000_000)), int(os.environ.get("AB_DIM", 128)), 10_000, 10
dev = torch.device("cuda:0")
data = gs.correlated_gpu(n, dim, dev, seed=42)
q = gs.subsample_with_noise_gpu(data, nq, seed=42)
lib = annb200.lib()
st = torch.cuda.current_stream(dev).cuda_stream
for name in dts:
    dt = {"f32": annb200.F32, "bf16": annb200.BF16, "sq8": annb200.SQ8}[name]
    ix = gs._flat_handle_from_device(data, metric, dt, 0)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ref = None
    for rep in range(2):
        for v in values:
            ix.set_option(opt, v)
            for _ in range(3):
                annb200._check(lib.annb_flat_search_dev(ix.handle, q.data_ptr(), nq, dim, k, ids.data_ptr(), d.data_ptr(), None, st))
            torch.cuda.synchronize()
            ix.set_option("time_kernels", 1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps = 20
            a.record()
            for _ in range(steps):
                annb200._check(lib.annb_flat_search_dev(ix.handle, q.data_ptr(), nq, dim, k, ids.data_ptr(), d.data_ptr(), None, st))
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / steps
            kern = ix.get_stat("dominant_kernel_ns") / max(1, ix.get_stat("dominant_kernel_launches")) * 1e-6
            same = True if ref is None else bool((ids == ref[0]).all() and (d.view(torch.int32) == ref[1].view(torch.int32)).all())
            if ref is None:
                ref = (ids.clone(), d.clone())
            print(f"flat {name} {opt} {v} rep {rep} step_ms {ms:.3f} kernel_ms {kern:.3f} qps {nq / ms * 1e3:.0f} same_as_first {same} "
                  f"uncertified {ix.get_stat('uncertified')} fallback_total {ix.get_stat('fallback_queries')} cert_eps {ix.cert_eps():.3e}", flush=True)
    ix.close()
