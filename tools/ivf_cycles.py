"""Prints the role wait-cycle counters of CTA 0 of the IVF tensor-core scan kernel (debug instrumentation).
usage: python tools/ivf_cycles.py [n] [nq] [nprobe] [dtype: f32|bf16|sq8] [world: the handle holds the first of `world` list shards]"""
import ctypes as C, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python"), os.path.join(ROOT, "tools")]
import annb200
import gpu_setup as gs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
nprobe = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dts = sys.argv[4].split(",") if len(sys.argv) > 4 else ["f32"]
world = int(sys.argv[5]) if len(sys.argv) > 5 else 1
dim, nlist, k = 128, 4096, 10
dev = torch.device("cuda:0")
data = gs.correlated_gpu(n, dim, dev, seed=42)
q = gs.subsample_with_noise_gpu(data, nq, seed=42)
lib = annb200.lib()
lib.annb_debug_fetch_cycles.argtypes = [C.c_void_p, C.c_void_p]
for name in dts:
    dt = {"f32": annb200.F32, "bf16": annb200.BF16, "sq8": annb200.SQ8}[name]
    parts = gs.build_ivf_parts_gpu(data, nlist, dt, 0, seed=42, kmeans_iters=8)
    if world > 1:
        from annb200 import distributed as D
        lb, le = D.list_ranges(parts["offsets"], world)[0]
        ix = gs.ivf_handle_from_parts(parts, n, dim, dt, annb200.L2, 0, lb, le)
    else:
        ix = gs.ivf_handle_from_parts(parts, n, dim, dt, annb200.L2, 0)
    ix.set_option("tc_debug", 1)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    for _ in range(3):
        annb200._check(lib.annb_ivf_search_dev(ix.handle, q.data_ptr(), nq, dim, k, nprobe, ids.data_ptr(), None, None, st))
    torch.cuda.synchronize()
    out = np.zeros(8, dtype=np.uint64)
    annb200._check(lib.annb_debug_fetch_cycles(ix.handle, out.ctypes.data_as(C.c_void_p)))
    tot, sched, gather, tfull, wq, wdata, wtempty, tt = [int(x) for x in out]
    tasks, tiles = tt >> 32, tt & 0xFFFFFFFF
    print(f"{name}: tasks={tasks} tiles={tiles} total={tot} cyc ({tot/1.965e6:.2f} ms) cyc/tile={tot/max(tiles,1):.0f} cyc/task={tot/max(tasks,1):.0f} | "
          f"schedule {sched/tot:.1%} ({sched/max(tasks,1):.0f}/task) | epilogue gather {gather/tot:.1%} ({gather/max(tasks,1):.0f}/task) wait-tfull {tfull/tot:.1%} | "
          f"mma wait-queries {wq/tot:.1%} wait-data {wdata/tot:.1%} wait-tempty {wtempty/tot:.1%}")
    ix.close()
