"""A/B of one integer index option (annb_index_set_option) on the 10M-row IVF bench workload, one index, one process.
usage: python tools/ivf_ab.py [n] [nq] [nprobe] [dtypes: f32,bf16,sq8] [option name] [values: 0,1,...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python"), os.path.join(ROOT, "tools")]
import annb200
import gpu_setup as gs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
nprobe = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dts = sys.argv[4].split(",") if len(sys.argv) > 4 else ["f32"]
opt = sys.argv[5] if len(sys.argv) > 5 else "ivf_list_major"
flag_sets = [int(x) for x in sys.argv[6].split(",")] if len(sys.argv) > 6 else [0, 1]
dim, nlist, k = 128, 4096, 10
dev = torch.device("cuda:0")
data = gs.correlated_gpu(n, dim, dev, seed=42)
q = gs.subsample_with_noise_gpu(data, nq, seed=42)
lib = annb200.lib()
st = torch.cuda.current_stream(dev).cuda_stream
for name in dts:
    dt = {"f32": annb200.F32, "bf16": annb200.BF16, "sq8": annb200.SQ8}[name]
    parts = gs.build_ivf_parts_gpu(data, nlist, dt, 0, seed=42, kmeans_iters=8)
    ix = gs.ivf_handle_from_parts(parts, n, dim, dt, annb200.L2, 0)
    for kv in os.environ.get("ANNB_AB_FIXED", "").split(","):      # options held fixed during the A/B, e.g. ivf_list_major=0
        if kv:
            ix.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    ix.query_batch(q.cpu().numpy(), k, nprobe=nprobe)              # host-buffer call: fills the probe statistics
    scanned_bytes = ix.get_stat("scanned_vectors") * dim * {"f32": 4, "bf16": 2, "sq8": 1}[name]
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    ref = None
    for rep in range(2):
        for flags in flag_sets:
            ix.set_option(opt, flags)
            for _ in range(3):
                annb200._check(lib.annb_ivf_search_dev(ix.handle, q.data_ptr(), nq, dim, k, nprobe, ids.data_ptr(), None, None, st))
            torch.cuda.synchronize()
            ix.set_option("time_kernels", 1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps = 20
            a.record()
            for _ in range(steps):
                annb200._check(lib.annb_ivf_search_dev(ix.handle, q.data_ptr(), nq, dim, k, nprobe, ids.data_ptr(), None, None, st))
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / steps
            kern = ix.get_stat("dominant_kernel_ns") / max(1, ix.get_stat("dominant_kernel_launches")) * 1e-6
            same = True if ref is None else bool((ids == ref).all())
            if ref is None:
                ref = ids.clone()
            print(f"{name} {opt} {flags} rep {rep} nq {nq} step_ms {ms:.3f} scan_ms {kern:.3f} qps {nq / ms * 1e3:.0f} ids_equal {same} "
                  f"algorithmic_GBs {scanned_bytes / (kern * 1e-3) / 1e9:.0f} path {ix.get_stat('last_path')} fallback_total {ix.get_stat('fallback_queries')}", flush=True)
    ix.close()
    del parts
    torch.cuda.empty_cache()
