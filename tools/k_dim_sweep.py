#!/usr/bin/env python
"""k / dim sweep of the flat f32 search on one GPU: tensor path (auto) against the exact CUDA-core path, device-resident
queries, CUDA events.  Shows where the tensor path's coverage ends (VERDICT r1 missing item 4: k > 24, dim > 128).

    python tools/k_dim_sweep.py [--n 1000000] [--nq 10000]
"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python"), os.path.join(ROOT, "tools")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--metric", default="cosine")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--simt", type=int, default=1)
    args = ap.parse_args()
    import torch
    import annb200
    import gpu_setup as gs
    L = annb200.lib()
    dev = torch.device("cuda", 0)
    metric = annb200.COSINE if args.metric == "cosine" else annb200.L2
    cases = [(128, k) for k in (10, 24, 25, 64, 128, 200, 256, 300)] + [(d, 10) for d in (64, 96, 160, 192, 256, 320, 512)] + [(256, 100)]
    print(f"# flat f32 {args.metric}, n={args.n}, {args.nq}-query batch, correlated synthetic; ms per batch (median of {args.steps})")
    print(f"{'dim':>4} {'k':>4} {'path':>7} {'tensor ms':>10} {'QPS':>10} {'fallback':>8} {'simt ms':>9} {'QPS':>9} {'speed-up':>8} {'ids==':>6}")
    last_dim, data, q, ix = None, None, None, None
    for dim, k in cases:
        if dim != last_dim:
            if ix is not None:
                ix.close()
            del data, q
            torch.cuda.empty_cache()
            data = gs.correlated_gpu(args.n, dim, dev, seed=42)
            q = gs.subsample_with_noise_gpu(data, args.nq, seed=42)
            ix = gs._flat_handle_from_device(data, metric, annb200.F32, 0)
            last_dim = dim
        st = torch.cuda.current_stream(dev).cuda_stream
        res = {}
        for path in (annb200.PATH_AUTO, annb200.PATH_SIMT):
            if path == annb200.PATH_SIMT and not args.simt:
                continue
            ix.set_option("path", path)
            ids = torch.empty((args.nq, k), dtype=torch.int64, device=dev)
            dd = torch.empty((args.nq, k), dtype=torch.float32, device=dev)
            f0 = ix.get_stat("fallback_queries")
            times = []
            steps = args.steps if path == annb200.PATH_AUTO else 2
            for it in range(1 + steps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                annb200._check(L.annb_flat_search_dev(ix.handle, q.data_ptr(), args.nq, dim, k, ids.data_ptr(), dd.data_ptr(), None, st))
                b.record()
                torch.cuda.synchronize()
                if it:
                    times.append(a.elapsed_time(b))
            res[path] = (float(np.median(times)), ix.get_stat("last_path"), (ix.get_stat("fallback_queries") - f0) // (1 + steps), ids.clone(), dd.clone())
        t = res[annb200.PATH_AUTO]
        line = f"{dim:>4} {k:>4} {'tensor' if t[1] == annb200.PATH_TENSOR else 'simt':>7} {t[0]:>10.3f} {args.nq / t[0] * 1e3:>10.0f} {t[2]:>8}"
        if annb200.PATH_SIMT in res:
            s = res[annb200.PATH_SIMT]
            same = bool(torch.equal(t[3], s[3]) and torch.equal(t[4].view(torch.int32), s[4].view(torch.int32)))
            line += f" {s[0]:>9.3f} {args.nq / s[0] * 1e3:>9.0f} {s[0] / t[0]:>8.2f} {str(same):>6}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
