#!/usr/bin/env python
"""bench.py -- headline measurement of the flat / IVF kNN hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload flat|ivf] ...
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one query batch.  Default workload (N = 1) is BASELINE.json
configs[1]: exhaustive flat f32 cosine, 1M x 128 Correlated synthetic, 10k-query batch, k = 10.
With N > 1 the database rows are sharded over the ranks (strong scaling: the database is fixed), every rank
searches its shard for the whole batch, per-shard top-k are all-gathered over NCCL and merged on the device.

Prints ONE JSON line (rank 0).  `value` = whole-job QPS with inputs resident in HBM; `e2e` = the same metric
through the host-buffer C-ABI call (pinned host queries in, host results out, copies inside the timed region).
`--impl reference` times the CPU restatement of the reference (oracle/, the one place besides cpu_baseline where
bench.py executes it) with all host threads on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="flat", choices=["flat", "ivf"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16", "sq8"])
    ap.add_argument("--metric", default=None, choices=[None, "cosine", "euclidean"])
    ap.add_argument("--n", "--rows", dest="n", type=int, default=None,
                    help="database rows (use --rows under torchrun: its own parser takes a bare --n for an abbreviation of --nnodes)")
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--nlist", type=int, default=4096)
    ap.add_argument("--nprobe", type=int, default=32)
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "tensor"])
    ap.add_argument("--cpu-sample", type=int, default=None, help="queries in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--recall", action="store_true", help="also report recall@k vs exact f32 ground truth on a query sample")
    ap.add_argument("--gpu-setup", default="auto", choices=["auto", "on", "off"],
                    help="IVF workload: generate data and build the index on the GPU (auto: when n > 2M)")
    ap.add_argument("--kmeans-iters", type=int, default=8)
    ap.add_argument("--tc-candidates", type=int, default=0, choices=[0, 16, 32], help="k' of the tensor-core pre-selection (0 = library default)")
    ap.add_argument("--replicated-routing", action="store_true", help="multi-GPU IVF: every rank ranks the centroids for the whole batch (no probe exchange)")
    ap.add_argument("--bf16-hybrid", type=int, default=-1, choices=[-1, 0, 1], help="flat bf16 tensor path: third query term in shared memory (-1 = library default)")
    ap.add_argument("--db-splits", type=int, default=0, help="flat tensor path: database splits per query tile (0 = library default)")
    ap.add_argument("--self-queries", action="store_true",
                    help="flat: the batch is database rows [0, nq) themselves (one batch of generate_knn; BASELINE configs[4]: --n 2000000 --dim 50 --k 15 --metric euclidean)")
    ap.add_argument("--list-major", type=int, default=-1, choices=[-1, 0, 1], help="IVF scan: -1 auto, 0 query-major streaming kernel, 1 list-major")
    ap.add_argument("--no-cert-fallback", action="store_true", help="diagnostic: do not read back / act on the uncertified count")
    ap.add_argument("--cert-eps-log2", type=int, default=0, help="log2 of the certificate's error bound (0 = library default)")
    return ap.parse_args()


# ----------------------------------------------------------------------------- helpers
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def make_data(args):
    from annb200 import datagen
    kind = "correlated"
    n = args.n
    data = datagen.make(kind, n, args.dim, seed=42)
    if args.self_queries:     # one batch of the all-vs-all kNN graph (exhaustive.rs:255-292): the rows themselves, self at rank 0
        queries = np.ascontiguousarray(data[:args.nq])
    else:
        queries = datagen.subsample_with_noise(data, args.nq, seed=42)
    return data, queries, kind


def metric_name(args):
    if args.workload == "flat":
        return f"QPS flat {args.dtype} {args.metric} k={args.k}" + (" self-query" if args.self_queries else "")
    return f"QPS ivf {args.dtype} {args.metric} nlist={args.nlist} nprobe={args.nprobe} k={args.k}"


def workload_desc(args, kind, n_gpus):
    if args.workload == "flat":
        w = f"exhaustive flat {args.dtype} {args.metric}, {args.n}x{args.dim} {kind} synthetic, {args.nq}-query batch, k={args.k}"
        if args.self_queries:
            w += f" (self-query: the batch is rows [0, {args.nq}) of the database; the full kNN graph is {-(-args.n // args.nq)} such batches)"
    else:
        w = (f"IVF {args.dtype} {args.metric}, {args.n}x{args.dim} {kind} synthetic, nlist={args.nlist}, nprobe={args.nprobe}, "
             f"{args.nq}-query batch, k={args.k}")
    return {"workload": w, "n": args.n, "dim": args.dim, "nq": args.nq, "k": args.k,
            "sharding": ("database rows" if args.workload == "flat" else "inverted lists") + f" over {n_gpus} GPU(s)" if n_gpus > 1 else "none",
            "l2_policy": "inputs larger than L2 (database streamed every step)"}


# ----------------------------------------------------------------------------- oracle-side (CPU) runs
def oracle_index(args, data):
    from oracle import oracle as o
    met = o.COSINE if args.metric == "cosine" else o.L2
    dt = {"f32": o.F32, "bf16": o.BF16, "sq8": o.SQ8}[args.dtype]
    if args.workload == "flat":
        return o.build_flat(data, met, dt)
    return o.build_ivf(data, met, nlist=args.nlist, dtype=dt, kmeans_iters=4)


def oracle_search(args, ix, q):
    from oracle import oracle as o
    if args.workload == "flat":
        return o.flat_search(ix, q, args.k)
    return o.ivf_search(ix, q, args.k, nprobe=args.nprobe)


def cpu_sample_size(args):
    if args.cpu_sample:
        return min(args.cpu_sample, args.nq)
    from oracle import oracle as o
    cores = o.max_threads()
    if args.workload == "flat":
        per_query_s = args.n * args.dim / 3.0e9            # ~3 G element-pairs / s / core (AVX2, memory bound)
    else:
        per_query_s = (args.nprobe * args.n / args.nlist + args.nlist) * args.dim / 3.0e9
    want_core_seconds = 16.0
    return int(max(cores, min(args.nq, want_core_seconds / max(per_query_s, 1e-9))))


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port: no Rust toolchain exists) on host cores."""
    from oracle import oracle as o
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    data, queries, kind = make_data(args)
    ix = oracle_index(args, data)
    ns = cpu_sample_size(args)
    q = queries[:ns]
    cores = o.max_threads()
    for _ in range(max(1, min(args.warmup, 1))):
        oracle_search(args, ix, q[:max(cores, 8)])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_search(args, ix, q)
    dt = (time.perf_counter() - t0) / args.steps
    qps = ns / dt
    line = {"impl": "reference", "metric": metric_name(args), "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_desc(args, kind, args.gpus),
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": f"{ns} of {args.nq} queries per step against the full database (linear in queries)"},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import annb200
    from annb200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    lib = annb200.lib()
    met = annb200.COSINE if args.metric == "cosine" else annb200.L2
    dt = {"f32": annb200.F32, "bf16": annb200.BF16, "sq8": annb200.SQ8}[args.dtype]
    path = {"auto": annb200.PATH_AUTO, "simt": annb200.PATH_SIMT, "tensor": annb200.PATH_TENSOR}[args.path]

    n, dim, nq, k = args.n, args.dim, args.nq, args.k
    algo_bytes_per_query = None
    gpu_setup = args.workload == "ivf" and (args.gpu_setup == "on" or (args.gpu_setup == "auto" and n > 2_000_000))
    truth_ids = None
    oi = None
    data = None
    if gpu_setup:
        # data + index built on the device with the library's own kernels (tools/gpu_setup.py); setup is not timed
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import gpu_setup as gs
        from annb200 import distributed as D
        kind = "correlated"
        t_setup = time.perf_counter()
        data_t = gs.correlated_gpu(n, dim, dev, seed=42)
        q_t = gs.subsample_with_noise_gpu(data_t, nq, seed=42)
        parts = gs.build_ivf_parts_gpu(data_t, args.nlist, dt, local_rank, seed=42, kmeans_iters=args.kmeans_iters)
        lb, le = D.list_ranges(parts["offsets"], world)[rank]
        index = gs.ivf_handle_from_parts(parts, n, dim, dt, met, local_rank, lb, le)
        queries = q_t.cpu().numpy()
        if rank == 0:
            truth_ids = gs.exact_ground_truth(data_t, q_t[:min(1000, nq)].contiguous(), k, met, local_rank)
            if not args.no_cpu_baseline:
                from oracle import oracle as o
                oi = o.IvfIndex({"f32": o.F32, "bf16": o.BF16, "sq8": o.SQ8}[args.dtype], o.L2 if met == annb200.L2 else o.COSINE, n, dim, args.nlist,
                                (parts["vectors"].view(torch.int16) if dt == annb200.BF16 else parts["vectors"]).cpu().numpy().view(
                                    {"f32": np.float32, "bf16": np.uint16, "sq8": np.int8}[args.dtype]),
                                parts["centroids"].cpu().numpy(), parts["offsets"].astype(np.int64), parts["original_ids"].cpu().numpy(),
                                scales=None if parts["scales"] is None else parts["scales"].cpu().numpy())
        del data_t, q_t, parts
        torch.cuda.empty_cache()
        if rank == 0:
            print(f"[setup] data + IVF index on GPU in {time.perf_counter() - t_setup:.1f} s", file=sys.stderr)
    else:
        data, queries, kind = make_data(args)
    if gpu_setup:
        pass
    elif args.workload == "flat":
        lo, hi = (rank * n) // world, ((rank + 1) * n) // world
        sq8_scales = annb200.sq8_train(annb200.normalise_rows(data) if met == annb200.COSINE else data) if dt == annb200.SQ8 and world > 1 else None
        index = annb200.ExhaustiveIndexB200.new(data[lo:hi], met, dt, device=local_rank, id_base=lo, sq8_scales=sq8_scales)
    else:
        from annb200 import distributed as D
        from oracle import oracle as o   # index *construction* for the small IVF bench uses the shared oracle build (setup, not timed)
        oi = oracle_index(args, data)
        lb, le = D.list_ranges(oi.offsets, world)[rank]
        r0, r1 = int(oi.offsets[lb]), int(oi.offsets[le])
        norms = oi.norms_i if oi.dtype == o.SQ8 else oi.norms
        index = annb200.IvfIndexB200.from_parts(oi.vectors[r0:r1], oi.centroids, oi.offsets, oi.original_ids[r0:r1], oi.dtype, oi.metric,
                                                norms=None if norms is None else norms[r0:r1], centroid_norms=oi.centroid_norms,
                                                sq8_scales=oi.scales, list_begin=lb, list_end=le, device=local_rank, n_total=oi.n)
    index.set_option("path", path)
    if args.tc_candidates:
        index.set_option("tc_candidates", args.tc_candidates)
    if args.cert_eps_log2:
        index.set_option("cert_eps_log2", args.cert_eps_log2)
    if args.db_splits:
        index.set_option("db_splits", args.db_splits)
    if args.bf16_hybrid >= 0:
        index.set_option("tc_bf16_hybrid", args.bf16_hybrid)
    if args.no_cert_fallback:
        index.set_option("cert_fallback", 0)
    if args.workload == "ivf" and args.list_major >= 0:
        index.set_option("ivf_list_major", args.list_major)
    index.set_option("time_kernels", 1)

    dq = torch.from_numpy(queries).to(dev)
    out_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    out_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    if world > 1:
        g_ids = torch.empty((world, nq, k), dtype=torch.int64, device=dev)
        g_dist = torch.empty((world, nq, k), dtype=torch.float32, device=dev)
        m_ids = torch.empty_like(out_ids)
        m_dist = torch.empty_like(out_dist)
    stream = torch.cuda.current_stream()
    merge_launches = 0

    def step_device():
        nonlocal merge_launches
        sp = stream.cuda_stream
        if args.workload == "flat":
            annb200._check(lib.annb_flat_search_dev(index.handle, dq.data_ptr(), nq, dim, k, out_ids.data_ptr(), out_dist.data_ptr(),
                                                    out_cnt.data_ptr(), sp))
        elif world > 1 and not args.replicated_routing:
            # every rank ranks the centroids for its slice of the batch only; probe lists are exchanged (annb200.distributed)
            D.ivf_search_sharded(index, dq, k, args.nprobe, m_ids, m_dist)
            merge_launches += 1
            return
        else:
            annb200._check(lib.annb_ivf_search_dev(index.handle, dq.data_ptr(), nq, dim, k, args.nprobe, out_ids.data_ptr(),
                                                   out_dist.data_ptr(), out_cnt.data_ptr(), sp))
        if world > 1:
            dist.all_gather_into_tensor(g_ids.view(-1), out_ids.view(-1))
            dist.all_gather_into_tensor(g_dist.view(-1), out_dist.view(-1))
            annb200._check(lib.annb_merge_topk_dev(g_ids.data_ptr(), g_dist.data_ptr(), world, nq, k, m_ids.data_ptr(), m_dist.data_ptr(),
                                                   None, sp))
            merge_launches += 1

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    sync_all()
    index.set_option("time_kernels", 1)          # reset the dominant-kernel accumulator after warm-up
    launches0 = index.get_stat("kernel_launches")
    merge_launches = 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    torch.cuda.profiler.start()     # `ncu --profile-from-start off` then sees exactly the timed region
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    sync_all()
    torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    dom_ns = index.get_stat("dominant_kernel_ns")
    dom_launches = max(1, index.get_stat("dominant_kernel_launches"))
    launches = index.get_stat("kernel_launches") - launches0 + merge_launches
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    qps = nq / (ms_per_step * 1e-3)
    last_path = index.get_stat("last_path")
    uncertified = index.get_stat("uncertified")      # tensor path, last step: queries recomputed on the exact path

    # ---- end to end through the host-buffer C ABI (pinned host queries, host outputs) ----
    hq = torch.from_numpy(queries).pin_memory()
    h_ids = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    h_dist = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    h_cnt = torch.empty((nq,), dtype=torch.int32).pin_memory()

    def step_host():
        if args.workload == "flat":
            annb200._check(lib.annb_flat_search(index.handle, hq.data_ptr(), nq, dim, k, h_ids.data_ptr(), h_dist.data_ptr(), h_cnt.data_ptr()))
        else:
            annb200._check(lib.annb_ivf_search(index.handle, hq.data_ptr(), nq, dim, k, args.nprobe, h_ids.data_ptr(), h_dist.data_ptr(),
                                               h_cnt.data_ptr()))
        if world > 1:   # shard results -> device -> all-gather -> merge -> host
            out_ids.copy_(h_ids, non_blocking=True)
            out_dist.copy_(h_dist, non_blocking=True)
            dist.all_gather_into_tensor(g_ids.view(-1), out_ids.view(-1))
            dist.all_gather_into_tensor(g_dist.view(-1), out_dist.view(-1))
            annb200._check(lib.annb_merge_topk_dev(g_ids.data_ptr(), g_dist.data_ptr(), world, nq, k, m_ids.data_ptr(), m_dist.data_ptr(),
                                                   None, stream.cuda_stream))
            h_ids.copy_(m_ids, non_blocking=True)
            h_dist.copy_(m_dist, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(max(1, args.warmup // 2)):
        step_host()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    sync_all()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_qps = nq / e2e_s
    # probed-list statistics are collected by the host-buffer entry point (last e2e step)
    scanned = index.get_stat("scanned_vectors") if args.workload == "ivf" else 0
    scanned_local = index.get_stat("scanned_vectors_local") if args.workload == "ivf" else 0   # this rank's own lists
    if world > 1 and args.workload == "ivf":
        # every rank derives the same global probe lists; the vectors it scans are those of its own lists
        pass

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peaks, peak_src = measured_peaks()
    dom_s = dom_ns * 1e-9 / dom_launches
    if args.workload == "flat":
        rows_local = (n + world - 1) // world
        flops = 2.0 * nq * rows_local * dim                      # algorithmic flops per launch: 2 * nq * n * d
        mma_per_elem = 3                                          # f32: 3xTF32 terms; bf16 index: f32 query = 3 bf16 terms
        if args.dtype == "f32":
            pipe_peak = peaks["bf16_tflops"] / 2.0               # TF32 pipe: measured bf16 / 2
            peak = pipe_peak / 3.0                               # BASELINE.md's f32 row: algorithmic flops executed 3x on the TF32 pipe
            peak_note = f"{peak_src} bf16 burst peak / 2 (TF32 pipe) / 3 (3xTF32 terms), as BASELINE.md section 2"
        elif args.dtype == "bf16":
            pipe_peak = peaks["bf16_tflops"]
            peak = pipe_peak                                     # BASELINE.md's bf16 row: one bf16 MMA per element
            peak_note = f"{peak_src} bf16 burst peak (the f32 query is fed as 3 bf16 terms, see 'executed')"
        else:
            mma_per_elem = 1                                     # int8 codes x int8 codes, s32 accumulate: one exact term
            pipe_peak = peaks["bf16_tflops"] * 2.0               # no measured int8 peak on this pool: nominal 2x the bf16 rate
            peak = pipe_peak
            peak_note = f"2 x {peak_src} bf16 burst peak (nominal int8:bf16 ratio; the int8 pipe is not measured separately)"
        if last_path != 2:
            mma_per_elem = 0
        achieved = flops / dom_s / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                    "kernel": "flat distance + top-k select", "kernel_ms": dom_s * 1e3, "peak_source": peak_note,
                    "executed": {"mma_terms_per_element": mma_per_elem, "tflops": achieved * mma_per_elem, "pipe_peak": pipe_peak,
                                 "frac_of_pipe_peak": achieved * mma_per_elem / pipe_peak},
                    "path": {0: "auto", 1: "simt (CUDA cores)", 2: "tensor (tcgen05)"}.get(last_path, str(last_path))}
    else:
        esz = {"f32": 4, "bf16": 2, "sq8": 1}[args.dtype]
        per_vec = dim * esz + (4 if args.metric == "cosine" else 0)
        algo_bytes = scanned_local * per_vec                     # sum over queries of probed list bytes on this rank, last step
        algo_bytes_per_query = scanned * per_vec / nq            # whole index
        hbm = {"algorithmic_bytes_per_launch": algo_bytes, "algorithmic_gbs": algo_bytes / dom_s / 1e9, "peak_gbs": peaks["hbm_gbs"]}
        if last_path == 2:
            # tensor-core grouped scan: one list load serves up to 128 queries, so the per-query byte count is not what the
            # kernel moves (algorithmic GB/s exceeds the HBM peak by design); the kernel is a grouped GEMM and is
            # bounded by the tensor pipe + its select epilogue.  HBM view: measured dram bytes from the ncu capture, if any.
            flops = 2.0 * scanned_local * dim                        # this rank's kernel scans its own lists only
            if args.dtype == "f32":
                pipe_peak, terms = peaks["bf16_tflops"] / 2.0, 3
            elif args.dtype == "bf16":
                pipe_peak, terms = peaks["bf16_tflops"], 3
            else:
                pipe_peak, terms = peaks["bf16_tflops"] * 2.0, 1
            peak = pipe_peak / (3.0 if args.dtype == "f32" else 1.0)
            achieved = flops / dom_s / 1e12
            roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                        "kernel": "ivf grouped list scan (tcgen05) + top-k' select", "kernel_ms": dom_s * 1e3,
                        "peak_source": f"{peak_src} bf16 burst peak scaled as for the flat kernel ({args.dtype})",
                        "algorithmic_flops_per_launch": flops,
                        "executed": {"mma_terms_per_element": terms, "note": "padded (list x 128-query group) tiles execute more MMA work than the algorithmic count"},
                        "hbm_view": hbm, "path": "tensor (tcgen05)"}
        else:
            achieved = algo_bytes / dom_s / 1e9
            peak = peaks["hbm_gbs"]
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                        "kernel": "ivf list scan (CUDA cores)", "kernel_ms": dom_s * 1e3, "peak_source": f"{peak_src} copy bandwidth",
                        "algorithmic_bytes_per_launch": algo_bytes, "path": "simt (CUDA cores)"}

    # DRAM traffic of the dominant kernel: from the committed ncu capture of this exact workload, if there is one
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        key = (f"flat {args.dtype} n={n} dim={dim} nq={nq}" if args.workload == "flat"
               else f"ivf {args.dtype} n={n} dim={dim} nq={nq} nlist={args.nlist} nprobe={args.nprobe}")
        if key in tr and world == 1:
            roofline["traffic"] = tr[key]["dram_bytes"]
            roofline["traffic_source"] = tr[key]["source"]
            if "hbm_view" in roofline:
                roofline["hbm_view"]["dram_gbs_from_traffic"] = tr[key]["dram_bytes"] / dom_s / 1e9
                roofline["hbm_view"]["frac_of_hbm_peak"] = tr[key]["dram_bytes"] / dom_s / 1e9 / peaks["hbm_gbs"]
    except Exception:
        pass
    line = {"metric": metric_name(args), "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"f32": "f32 (3xTF32 select + f32 exact re-rank)" if last_path == 2 else "f32", "bf16": "bf16 (f32 accumulate)",
                      "sq8": "int8 (i32 accumulate)"}[args.dtype],
            "data": "synthetic", "config": workload_desc(args, kind, world), "clocks": clocks,
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": int(nq * dim * 4),
                    "d2h_bytes_per_step": int(nq * k * 12 + nq * 4), "ms_per_step": e2e_s * 1e3},
            "gpu_launches": int(launches), "roofline": roofline, "uncertified_queries_last_step": int(uncertified)}
    if algo_bytes_per_query is not None:
        line["config"]["algorithmic_bytes_per_query"] = algo_bytes_per_query
    if args.self_queries:
        line["full_knn_graph_seconds_extrapolated"] = n / qps

    # ---- parity spot check + recall on a query sample (outside the timed region) ----
    from oracle import oracle as o
    got_ids_all = h_ids.numpy()
    got_d_all = h_dist.numpy()
    if world > 1:   # the device-resident leg and the host-buffer leg are different call chains: their merged results must agree
        line["device_vs_e2e"] = {"dist_bits_equal": bool(np.array_equal(m_dist.cpu().numpy().view(np.uint32), got_d_all.view(np.uint32)))}
    if world == 1 and (args.workload == "flat" or oi is not None):
        ns = min(64, nq)
        oi2 = oracle_index(args, data) if args.workload == "flat" else oi
        ref = oracle_search(args, oi2, queries[:ns])
        line["parity_sample"] = {"queries": ns, "ids_equal": bool(np.array_equal(got_ids_all[:ns], ref[0])),
                                 "dist_bits_equal": bool(np.array_equal(got_d_all[:ns].view(np.uint32), ref[1].view(np.uint32)))}
    if truth_ids is not None:
        line["recall_at_k_vs_exact_f32"] = {"value": o.recall_at_k(truth_ids, got_ids_all[:truth_ids.shape[0]], k), "queries": int(truth_ids.shape[0])}
    elif data is not None and (args.workload == "ivf" or args.dtype != "f32"):
        ns = min(64, nq)
        exact = o.flat_search(o.build_flat(data, o.COSINE if args.metric == "cosine" else o.L2), queries[:ns], k)
        line["recall_at_k_vs_exact_f32"] = {"value": o.recall_at_k(exact[0], got_ids_all[:ns], k), "queries": ns}

    # ---- CPU baseline on this host's cores (bounded sample) ----
    if not args.no_cpu_baseline and (args.workload == "flat" or oi is not None):
        oi3 = oracle_index(args, data) if args.workload == "flat" else oi
        ns = cpu_sample_size(args)
        cores = o.max_threads()
        oracle_search(args, oi3, queries[:max(8, min(ns, cores))])
        t0 = time.perf_counter()
        oracle_search(args, oi3, queries[:ns])
        dtc = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ns / dtc, "unit": "queries/s", "cores": cores, "kind": "port",
                                "sample": f"{ns} of {nq} queries against the full database, {dtc:.2f} s wall (QPS is linear in queries)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.metric is None:
        args.metric = "cosine" if args.workload == "flat" else "euclidean"
    if args.n is None:
        args.n = 1_000_000 if args.workload == "flat" else 10_000_000
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
