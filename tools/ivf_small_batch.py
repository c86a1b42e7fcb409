"""IVF scan kernel choice for small query batches: times annb_ivf_search_dev with the query-major streaming kernel
(ivf_list_major = 0), the list-major batched kernel (1) and the library's own choice (-1) over a sweep of batch sizes.
usage: python tools/ivf_small_batch.py [n] [nprobe] [dtype: f32|bf16|sq8] [nq,nq,...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python"), os.path.join(ROOT, "tools")]
import annb200
import gpu_setup as gs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nprobe = int(sys.argv[2]) if len(sys.argv) > 2 else 32
name = sys.argv[3] if len(sys.argv) > 3 else "f32"
nqs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 4, 8, 16, 32, 64, 128, 256, 1000]
dim, nlist, k = 128, 4096, 10
dev = torch.device("cuda:0")
data = gs.correlated_gpu(n, dim, dev, seed=42)
q_all = gs.subsample_with_noise_gpu(data, max(nqs), seed=42)
lib = annb200.lib()
dt = {"f32": annb200.F32, "bf16": annb200.BF16, "sq8": annb200.SQ8}[name]
parts = gs.build_ivf_parts_gpu(data, nlist, dt, 0, seed=42, kmeans_iters=8)
ix = gs.ivf_handle_from_parts(parts, n, dim, dt, annb200.L2, 0)
st = torch.cuda.current_stream(dev).cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for nq in nqs:
    q = q_all[:nq].contiguous()
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    ref = None
    for mode in (0, 1, -1):
        ix.set_option("ivf_list_major", mode)
        for _ in range(3):
            annb200._check(lib.annb_ivf_search_dev(ix.handle, q.data_ptr(), nq, dim, k, nprobe, ids.data_ptr(), None, None, st))
        reps, tot = 10, 0.0
        for _ in range(reps):
            flush.zero_()   # the probed lists of a tiny batch would otherwise stay in L2
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            annb200._check(lib.annb_ivf_search_dev(ix.handle, q.data_ptr(), nq, dim, k, nprobe, ids.data_ptr(), None, None, st))
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        same = True if ref is None else bool((ids == ref).all())
        if ref is None:
            ref = ids.clone()
        print(f"nq {nq} list_major {mode} ms {tot / reps:.3f} qps {nq / (tot / reps) * 1e3:.0f} ids_equal_to_mode0 {same}", flush=True)
ix.close()
