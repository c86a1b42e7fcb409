#!/usr/bin/env python
"""bench.py -- measurement of the flat / IVF kNN hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload all|flat|ivf|c5] ...
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one query batch.  The default run (`--workload all`) prints ONE JSON line whose
top-level fields are the headline, BASELINE.json configs[1] -- exhaustive flat f32 cosine, 1M x 128 Correlated synthetic,
10k-query batch, k = 10 -- and which carries two more blocks measured the same way in the same process:

    "ivf": BASELINE configs[2] and [3]: IVF 10M x 128, nlist 4096, L2, k = 10: f32 at nprobe 8 / 32 / 128, BF16 and SQ8 at
           nprobe 32 (one shared build, on the device); each entry has value / e2e / roofline (tensor view AND HBM view) /
           recall@10 vs exact f32 ground truth / a parity sample against the CPU oracle on the same index contents /
           the CPU port timed on this host's cores;
    "c5":  BASELINE configs[4]: one 10k-row batch of the all-vs-all kNN graph, 2M x 50, k = 15.

With N > 1 (one process per GPU) database rows (flat, c5) or inverted lists (IVF) are sharded over the ranks (strong
scaling: the database is fixed); per-shard top-k travel in ONE all-gather over NCCL and are merged on the device
(annb200.distributed.ShardedSearch).  `--single-process` instead drives all N GPUs from this one process through the
library's own multi-device handle (annb_*_create_multi: peer copies over NVLink, no NCCL).

`value` = whole-job QPS with inputs resident in HBM; `e2e` = the same metric through the host-buffer call (pinned host
queries in, host results out, copies inside the timed region).  `--impl reference` times the CPU restatement of the
reference (oracle/, the one place besides cpu_baseline / parity samples where bench.py executes it) with all host
threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "flat", "ivf", "c5"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16", "sq8"])
    ap.add_argument("--metric", default=None, choices=[None, "cosine", "euclidean"])
    ap.add_argument("--n", "--rows", dest="n", type=int, default=None,
                    help="database rows (use --rows under torchrun: its own parser takes a bare --n for an abbreviation of --nnodes)")
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--nlist", type=int, default=4096)
    ap.add_argument("--nprobe", type=int, default=32)
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "tensor"])
    ap.add_argument("--cpu-sample", type=int, default=None, help="queries in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kmeans-iters", type=int, default=8)
    ap.add_argument("--tc-candidates", type=int, default=0, choices=[0, 16, 32], help="k' of the tensor-core pre-selection (0 = library default)")
    ap.add_argument("--db-splits", type=int, default=0, help="flat tensor path: database splits per query tile (0 = library default)")
    ap.add_argument("--self-queries", action="store_true",
                    help="flat: the batch is database rows [0, nq) themselves (one batch of generate_knn)")
    ap.add_argument("--list-major", type=int, default=-1, choices=[-1, 0, 1], help="IVF scan: -1 auto, 0 query-major streaming kernel, 1 list-major")
    ap.add_argument("--no-cert-fallback", action="store_true", help="diagnostic: do not act on the uncertified count")
    ap.add_argument("--cert-eps-log2", type=int, default=1, help="log2 of the certificate's error bound (1 = derived bound, the library default; 0 = off)")
    ap.add_argument("--single-process", action="store_true", help="drive --gpus N devices from this one process through annb_*_create_multi")
    ap.add_argument("--ivf-set", default="f32:8,f32:32,f32:128,bf16:32,sq8:32", help="--workload all block: dtype:nprobe entries")
    ap.add_argument("--option", action="append", default=[], help="extra index option key=value (repeatable)")
    return ap.parse_args()


# ----------------------------------------------------------------------------- helpers
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed regions (B200_PROFILING.md recipe)."""

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
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        """Median SM clock and throttle reasons of the samples taken in [t0, t1] (falls back to the nearest samples)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if t1 - t0 < 0.25:
            time.sleep(0.25)            # short regions: make sure at least one sample taken right behind the region has arrived
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.12] or [r for (_, r) in self.rows[-3:]]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if not self.proc:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "MEASURED_PEAKS.json"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback of B200_PROFILING.md"


def traffic_table():
    out = {}
    for name in ("r01_traffic.json", "r02_traffic.json"):
        try:
            out.update(json.load(open(os.path.join(ROOT, "profiles", name))))
        except Exception:
            pass
    return out


# ----------------------------------------------------------------------------- reference arm (CPU port)
def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port: no Rust toolchain exists) on ALL host cores.
    Under torchrun the launcher exports OMP_NUM_THREADS=1; the thread count is therefore passed explicitly."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from annb200 import datagen
    from oracle import oracle as o
    cores = os.cpu_count() or 1
    n, dim, nq, k = args.n or 1_000_000, args.dim or 128, args.nq, args.k or 10
    metric = args.metric or "cosine"
    data = datagen.make("correlated", n, dim, seed=42)
    queries = datagen.subsample_with_noise(data, nq, seed=42)
    met = o.COSINE if metric == "cosine" else o.L2
    dt = {"f32": o.F32, "bf16": o.BF16, "sq8": o.SQ8}[args.dtype]
    ix = o.build_flat(data, met, dt)
    per_query_s = n * dim / 3.0e9                       # ~3 G element-pairs / s / core (AVX2, memory bound)
    ns = args.cpu_sample or int(max(cores, min(nq, 1.0 * cores / max(per_query_s, 1e-9))))   # ~1 s of wall per step
    ns = min(ns, nq)
    q = queries[:ns]
    for _ in range(max(1, min(args.warmup, 2))):
        o.flat_search(ix, q[:max(cores, 8)], k, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.flat_search(ix, q, k, nthreads=cores)
    dt_s = (time.perf_counter() - t0) / args.steps
    qps = ns / dt_s
    line = {"impl": "reference", "metric": f"QPS flat {args.dtype} {metric} k={k}", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_s * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"exhaustive flat {args.dtype} {metric}, {n}x{dim} correlated synthetic, {nq}-query batch, k={k}", "n": n, "dim": dim,
                       "nq": nq, "k": k},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": f"{ns} of {nq} queries per step against the full database (linear in queries); {cores} OpenMP threads "
                                       f"(OMP_NUM_THREADS in the environment was {os.environ.get('OMP_NUM_THREADS', 'unset')}, overridden)"},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- B200 arm
class Ctx:
    pass


def init_ctx(args):
    import torch
    import torch.distributed as dist

    import annb200
    c = Ctx()
    c.torch, c.dist, c.annb = torch, dist, annb200
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if c.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's banner off stdout: rank 0 prints exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{c.local_rank}"))
    torch.cuda.set_device(c.local_rank)
    c.dev = torch.device(f"cuda:{c.local_rank}")
    c.lib = annb200.lib()
    c.single = args.single_process and args.gpus > 1 and c.world == 1
    c.devices = list(range(args.gpus)) if c.single else [c.local_rank]
    c.shards = args.gpus if c.single else c.world          # number of GPUs the index is spread over
    c.peaks, c.peak_src = measured_peaks()
    c.traffic = traffic_table()
    c.sampler = ClockSampler(c.local_rank)
    if c.rank == 0:
        c.sampler.start()
    c.args = args
    return c


def measure_tf32_peak(c):
    """cuBLAS TF32 GEMM rate on this GPU, measured in-run: the denominator of the f32 (3xTF32) roofline rows
    (MEASURED_PEAKS.json carries the bf16 rate only).  A plain library GEMM, used for nothing but this number."""
    torch = c.torch
    try:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn((8192, 8192), device=c.dev)
        b = torch.randn((8192, 8192), device=c.dev)
        best = None
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if i >= 2:
                best = ms if best is None else min(best, ms)
        torch.backends.cuda.matmul.allow_tf32 = prev
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def apply_options(args, index):
    annb200 = sys.modules["annb200"]
    path = {"auto": annb200.PATH_AUTO, "simt": annb200.PATH_SIMT, "tensor": annb200.PATH_TENSOR}[args.path]
    index.set_option("path", path)
    if args.tc_candidates:
        index.set_option("tc_candidates", args.tc_candidates)
    if args.cert_eps_log2 != 1:
        index.set_option("cert_eps_log2", args.cert_eps_log2)
    if args.db_splits:
        index.set_option("db_splits", args.db_splits)
    if args.no_cert_fallback:
        index.set_option("cert_fallback", 0)
    if args.list_major >= 0 and index.info().is_ivf:
        index.set_option("ivf_list_major", args.list_major)
    for kv in args.option:
        key, val = kv.split("=")
        index.set_option(key, int(val))


def timed_run(c, index, step_device, step_host, nq, steps, warmup, flush=None):
    """W untimed warm-up steps, then exactly `steps` steps bracketed by barrier + synchronize on both sides, CUDA events on
    the launching stream, max over ranks; then the same for the host-buffer (end-to-end) step, wall clock.
    flush: finishes steps the device loop left pending (sharded runs defer every step's verdict by one step); it runs INSIDE
    the timed region, so a refine any verdict asks for is paid for there."""
    torch, dist = c.torch, c.dist
    flush = flush or (lambda: None)

    def sync_all():
        torch.cuda.synchronize()
        if c.single:
            for d in c.devices:
                torch.cuda.synchronize(d)
        if c.world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    for _ in range(warmup):
        step_device()
    flush()
    sync_all()
    index.set_option("time_kernels", 1)          # resets the dominant-kernel accumulator after warm-up
    launches0 = index.get_stat("kernel_launches")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_w0 = time.perf_counter()
    torch.cuda.profiler.start()     # `ncu --profile-from-start off` then sees exactly the timed region
    e0.record(stream)
    for _ in range(steps):
        step_device()
    flush()
    e1.record(stream)
    sync_all()
    torch.cuda.profiler.stop()
    t_w1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    if c.single:
        ms = (t_w1 - t_w0) * 1e3      # several devices, one process: every step ends with a host-side join, wall clock brackets it
    dom_ns = index.get_stat("dominant_kernel_ns")
    dom_launches = max(1, index.get_stat("dominant_kernel_launches"))
    launches = index.get_stat("kernel_launches") - launches0
    index.set_option("time_kernels", 0)
    if c.world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=c.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    for _ in range(max(1, warmup // 2)):
        step_host()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_host()
    sync_all()
    e2e_s = (time.perf_counter() - t0) / steps
    if c.world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=c.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    return {"ms_per_step": ms / steps, "qps": nq / (ms / steps * 1e-3), "e2e_s": e2e_s, "e2e_qps": nq / e2e_s, "dom_s": dom_ns * 1e-9 / dom_launches,
            "launches": int(launches), "clocks": c.sampler.window(t_w0, t_w1) if c.rank == 0 else None}


class Searcher:
    """The device-resident step and the host-buffer (end-to-end) step of one index, for 1 GPU, N processes or N devices in one process."""

    def __init__(self, c, index, ivf, nq, dim, k, nprobe, queries_np, self_rows=None):
        torch = c.torch
        self.c, self.index, self.ivf, self.nq, self.dim, self.k, self.nprobe = c, index, ivf, nq, dim, k, nprobe
        self.dq = torch.from_numpy(queries_np).to(c.dev)
        self.hq = torch.from_numpy(queries_np).pin_memory()
        self.out_ids = torch.empty((nq, k), dtype=torch.int64, device=c.dev)
        self.out_dist = torch.empty((nq, k), dtype=torch.float32, device=c.dev)
        self.h_ids = torch.empty((nq, k), dtype=torch.int64).pin_memory()
        self.h_dist = torch.empty((nq, k), dtype=torch.float32).pin_memory()
        self.h_cnt = torch.empty((nq,), dtype=torch.int32).pin_memory()
        self.self_rows = self_rows                     # (begin, end): the batch is these resident rows (1 GPU / single process)
        self.extra_launches = 0
        self.sharded = None
        self.pair = None
        self.turn = 0
        if c.world > 1:
            from annb200 import distributed as D
            # two step objects (each with its own exchange buffers) take turns in the device loop: step i + 1 is enqueued before
            # the verdict of step i is read, so the host runs one step ahead of the devices (ShardedSearch, defer=True)
            self.pair = [D.ShardedSearch(index, nq, dim, k, nprobe if ivf else 0, None, c.dev) for _ in range(2)]
            self.sharded = self.pair[0]

    def flush(self):
        if self.pair is not None:
            for s in self.pair:
                s.resolve()

    def step_device(self):
        c, lib, annb200 = self.c, self.c.lib, self.c.annb
        sp = c.torch.cuda.current_stream().cuda_stream
        if self.pair is not None:
            s = self.pair[self.turn & 1]
            self.turn += 1
            s(self.dq, defer=True)                     # (resolves its own previous step first)
            self.sharded = s                           # device_result() reads the last step's object
            self.extra_launches += 2                   # the merge and check kernels
            return
        if self.ivf:
            annb200._check(lib.annb_ivf_search_dev(self.index.handle, self.dq.data_ptr(), self.nq, self.dim, self.k, self.nprobe, self.out_ids.data_ptr(),
                                                   self.out_dist.data_ptr(), None, sp))
        else:
            annb200._check(lib.annb_flat_search_dev(self.index.handle, self.dq.data_ptr(), self.nq, self.dim, self.k, self.out_ids.data_ptr(),
                                                    self.out_dist.data_ptr(), None, sp))

    def step_host(self):
        c, lib, annb200 = self.c, self.c.lib, self.c.annb
        if self.sharded is not None:
            # one process per GPU: pinned host queries -> device, sharded step, merged result -> pinned host (no bounce through the host in between)
            self.flush()
            self.dq.copy_(self.hq, non_blocking=True)
            ids, dst = self.sharded(self.dq)
            self.h_ids.copy_(ids, non_blocking=True)
            self.h_dist.copy_(dst, non_blocking=True)
            c.torch.cuda.synchronize()
            return
        if self.self_rows is not None and not self.ivf:
            annb200._check(lib.annb_flat_search_self(self.index.handle, self.self_rows[0], self.self_rows[1], self.k, self.h_ids.data_ptr(),
                                                     self.h_dist.data_ptr(), self.h_cnt.data_ptr()))
        elif self.ivf:
            annb200._check(lib.annb_ivf_search(self.index.handle, self.hq.data_ptr(), self.nq, self.dim, self.k, self.nprobe, self.h_ids.data_ptr(),
                                               self.h_dist.data_ptr(), self.h_cnt.data_ptr()))
        else:
            annb200._check(lib.annb_flat_search(self.index.handle, self.hq.data_ptr(), self.nq, self.dim, self.k, self.h_ids.data_ptr(),
                                                self.h_dist.data_ptr(), self.h_cnt.data_ptr()))

    def device_result(self):
        self.flush()
        if self.sharded is not None:
            return self.sharded.out_ids.cpu().numpy(), self.sharded.out_dist.cpu().numpy()
        return self.out_ids.cpu().numpy(), self.out_dist.cpu().numpy()


def tensor_roofline(c, dtype, flops, dom_s, kernel, tf32_peak, bf16_terms=2, f32_kind=0):
    """Roofline row of a tensor-core kernel: achieved = algorithmic flops / kernel time.
    bf16_terms: bf16 terms the f32 query is fed as (library default: 2 for cosine, 3 for squared Euclidean).
    f32_kind: operand form of the f32 kernel (stat tc_kind): 0 = 3xTF32 (kind::tf32 pipe), 3 = 3xFP16 (kind::f16 pipe: the same three
    product terms at 16 elements per MMA K step, so its peak is the 16-bit pipe's, not the TF32 pipe's)."""
    peaks = c.peaks
    if dtype == "f32" and f32_kind == 3:
        pipe_peak = pipe_alt = peaks["bf16_tflops"]
        peak, terms = pipe_peak / 3.0, 3
        note = (f"{c.peak_src} 16-bit tensor burst peak ({pipe_peak:.1f} TFLOP/s, cuBLAS bf16) / 3 (3xFP16 terms); "
                f"sustained figure of the same file: {peaks.get('bf16_tflops_sustained', 0):.1f}")
    elif dtype == "f32":
        pipe_alt = peaks["bf16_tflops"] / 2.0
        pipe_peak = tf32_peak if tf32_peak else pipe_alt
        peak = pipe_peak / 3.0
        terms = 3
        note = (f"cuBLAS TF32 GEMM measured in this run ({pipe_peak:.1f} TFLOP/s) / 3 (3xTF32 terms)" if tf32_peak else
                f"{c.peak_src} bf16 burst peak / 2 (TF32 pipe) / 3 (3xTF32 terms)")
    elif dtype == "bf16":
        pipe_peak = pipe_alt = peaks["bf16_tflops"]
        peak, terms = pipe_peak, bf16_terms
        note = f"{c.peak_src} bf16 burst peak (the f32 query is fed as {bf16_terms} bf16 terms, see 'executed')"
    else:
        pipe_peak = pipe_alt = peaks["bf16_tflops"] * 2.0
        peak, terms = pipe_peak, 1
        note = f"2 x {c.peak_src} bf16 burst peak (nominal int8:bf16 ratio; the int8 pipe is not measured separately)"
    achieved = flops / dom_s / 1e12 if dom_s > 0 else 0.0
    out = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None, "kernel": kernel,
           "kernel_ms": dom_s * 1e3, "peak_source": note, "pipe_peak_from_bf16": pipe_alt,
           "executed": {"mma_terms_per_element": terms, "tflops": achieved * terms, "pipe_peak": pipe_peak, "frac_of_pipe_peak": achieved * terms / pipe_peak}}
    if dtype == "f32" and f32_kind == 3 and peaks.get("bf16_tflops_sustained"):
        out["executed"]["frac_of_sustained_pipe_peak"] = achieved * terms / peaks["bf16_tflops_sustained"]
        out["executed"]["operand_form"] = "3xFP16: rows scaled by powers of two, fp16 hi + lo pieces, f32 accumulate"
    return out


def tie_classes_equal(ids, dist, ref_ids, ref_dist):
    """Same rule as tests/util.py::assert_tie_classes: distance bits equal, and inside every run of equal distances that ends
    before the last valid slot the id SETS agree (the reference's IVF-SQ8 heap, src/quantised/ivf_sq8.rs:329-352, keeps an
    implementation-defined subset of the last tie class; integer distances tie often)."""
    if not np.array_equal(np.asarray(dist).view(np.uint32), np.asarray(ref_dist).view(np.uint32)):
        return False
    for r in range(ids.shape[0]):
        valid = int((ids[r] >= 0).sum())
        if valid != int((ref_ids[r] >= 0).sum()):
            return False
        j = 0
        while j < valid:
            e = j
            while e + 1 < valid and dist[r, e + 1] == dist[r, j]:
                e += 1
            if e < valid - 1:
                if set(ids[r, j:e + 1].tolist()) != set(ref_ids[r, j:e + 1].tolist()):
                    return False
            elif len(set(ids[r, j:e + 1].tolist())) != e + 1 - j:
                return False
            j = e + 1
    return True


def cpu_baseline_flat(oix, queries, k, n, dim, nq, self_mode=False):
    from oracle import oracle as o
    cores = os.cpu_count() or 1
    per_query_s = n * dim / 3.0e9
    ns = int(max(cores, min(nq, 16.0 / max(per_query_s, 1e-9))))       # ~16 core-seconds
    if self_mode:
        rows = np.arange(ns, dtype=np.int64)
        o.flat_search(oix, None, k, self_rows=rows[:max(8, min(ns, cores))], self_mode=True, nthreads=cores)
        t0 = time.perf_counter()
        o.flat_search(oix, None, k, self_rows=rows, self_mode=True, nthreads=cores)
    else:
        o.flat_search(oix, queries[:max(8, min(ns, cores))], k, nthreads=cores)
        t0 = time.perf_counter()
        o.flat_search(oix, queries[:ns], k, nthreads=cores)
    dtc = time.perf_counter() - t0
    return {"value": ns / dtc, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{ns} of {nq} queries against the full database, {dtc:.2f} s wall (QPS is linear in queries)"}


# ---- flat (headline, and c5 when self_queries) --------------------------------------------------------------------
def bench_flat(c, n, dim, nq, k, metric, dtype, self_queries, tf32_peak, on_gpu_data=False):
    torch, annb200, args = c.torch, c.annb, c.args
    from annb200 import datagen
    from oracle import oracle as o
    met = annb200.COSINE if metric == "cosine" else annb200.L2
    dt = {"f32": annb200.F32, "bf16": annb200.BF16, "sq8": annb200.SQ8}[dtype]
    kind = "correlated"
    if on_gpu_data:
        import gpu_setup as gs
        data_t = gs.correlated_gpu(n, dim, c.dev, seed=42)
        data = data_t.cpu().numpy()
        del data_t
        torch.cuda.empty_cache()
    else:
        data = datagen.make(kind, n, dim, seed=42)
    queries = np.ascontiguousarray(data[:nq]) if self_queries else datagen.subsample_with_noise(data, nq, seed=42)
    if c.single:
        index = annb200.ExhaustiveIndexB200.new(data, met, dt, device=c.devices)
    else:
        lo, hi = (c.rank * n) // c.world, ((c.rank + 1) * n) // c.world
        sq8_scales = annb200.sq8_train(annb200.normalise_rows(data) if met == annb200.COSINE else data) if dt == annb200.SQ8 and c.world > 1 else None
        index = annb200.ExhaustiveIndexB200.new(data[lo:hi], met, dt, device=c.local_rank, id_base=lo, sq8_scales=sq8_scales)
    apply_options(args, index)
    se = Searcher(c, index, False, nq, dim, k, 0, queries, self_rows=(0, nq) if (self_queries and c.world == 1) else None)
    r = timed_run(c, index, se.step_device, se.step_host, nq, args.steps, args.warmup, se.flush)
    last_path = index.get_stat("last_path")
    uncertified = index.get_stat("uncertified")
    fallback_q = index.get_stat("fallback_queries")
    dev_ids, dev_dist = se.device_result()
    if c.rank != 0:
        index.close()
        return None
    rows_local = (n + c.shards - 1) // c.shards
    flops = 2.0 * nq * rows_local * dim                          # algorithmic flops per launch: 2 * nq * n_local * d
    try:
        f32_kind = index.get_stat("tc_kind")
    except Exception:
        f32_kind = 0
    roofline = tensor_roofline(c, dtype, flops, r["dom_s"], "flat distance + top-k select (flat_tc_kernel)", tf32_peak,
                               bf16_terms=(1 if self_queries else (2 if metric == "cosine" else 3)), f32_kind=f32_kind)
    roofline["path"] = {0: "auto", 1: "simt (CUDA cores)", 2: "tensor (tcgen05)"}.get(last_path, str(last_path))
    if last_path != 2:
        roofline["executed"]["mma_terms_per_element"] = 0
    key = f"flat {dtype} n={n} dim={dim} nq={nq}" + (" self" if self_queries else "")
    if key in c.traffic and c.shards == 1:
        roofline["traffic"] = c.traffic[key]["dram_bytes"]
        roofline["traffic_source"] = c.traffic[key]["source"]
    w = f"exhaustive flat {dtype} {metric}, {n}x{dim} {kind} synthetic, {nq}-query batch, k={k}"
    if self_queries:
        w += f" (self-query: the batch is rows [0, {nq}) of the database; the full kNN graph is {-(-n // nq)} such batches)"
    line = {"metric": f"QPS flat {dtype} {metric} k={k}" + (" self-query" if self_queries else ""), "value": r["qps"], "unit": "queries/s",
            "n_gpus": c.shards, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": {"f32": ("f32 (3xFP16 split-precision select + f32 exact re-rank)" if f32_kind == 3 else "f32 (3xTF32 select + f32 exact re-rank)") if last_path == 2 else "f32", "bf16": "bf16 (f32 accumulate)",
                      "sq8": "int8 (i32 accumulate)"}[dtype],
            "data": "synthetic",
            "config": {"workload": w, "n": n, "dim": dim, "nq": nq, "k": k,
                       "sharding": (f"database rows over {c.shards} GPUs, " + ("one process, peer copies" if c.single else "one process per GPU, one NCCL all-gather; device loop: verdict of step i read after step i+1 is enqueued"))
                       if c.shards > 1 else "none",
                       "l2_policy": "inputs larger than L2 (database streamed every step)"},
            "clocks": r["clocks"],
            "e2e": {"value": r["e2e_qps"], "unit": "queries/s", "h2d_bytes_per_step": 0 if (self_queries and c.world == 1) else int(nq * dim * 4),
                    "d2h_bytes_per_step": int(nq * k * 12 + (nq * 4 if c.world == 1 else 0)), "ms_per_step": r["e2e_s"] * 1e3},
            "gpu_launches": r["launches"] + se.extra_launches, "roofline": roofline, "uncertified_queries_last_step": int(uncertified),
            "fallback_queries_total": int(fallback_q)}
    if se.pair is not None:
        line["shard_refined_queries_total"] = int(sum(s.refined_queries for s in se.pair))
    if self_queries:
        line["full_knn_graph_seconds_extrapolated"] = n / r["qps"]
    # parity sample against the CPU oracle, recall for the quantised indices, CPU baseline
    got_ids, got_d = se.h_ids.numpy(), se.h_dist.numpy()
    if c.shards > 1:   # device-resident leg and host-buffer leg are different call chains: both must give the same bits
        line["device_vs_e2e"] = {"ids_equal": bool(np.array_equal(dev_ids, got_ids)),
                                 "dist_bits_equal": bool(np.array_equal(dev_dist.view(np.uint32), got_d.view(np.uint32)))}
    ns = min(64, nq)
    cores = os.cpu_count() or 1
    oix = o.build_flat(data, o.COSINE if metric == "cosine" else o.L2, {"f32": o.F32, "bf16": o.BF16, "sq8": o.SQ8}[dtype])
    if self_queries:
        ref = o.flat_search(oix, None, k, self_rows=np.arange(ns, dtype=np.int64), self_mode=True, nthreads=cores)
    else:
        ref = o.flat_search(oix, queries[:ns], k, nthreads=cores)
    line["parity_sample"] = {"queries": ns, "ids_equal": bool(np.array_equal(got_ids[:ns], ref[0])),
                             "dist_bits_equal": bool(np.array_equal(got_d[:ns].view(np.uint32), ref[1].view(np.uint32))),
                             "against": "CPU oracle (oracle/oracle.c), same data" + (f", merged over {c.shards} shards" if c.shards > 1 else "")}
    if dtype != "f32":
        exact = o.flat_search(o.build_flat(data, o.COSINE if metric == "cosine" else o.L2), queries[:ns], k, nthreads=cores)
        line["recall_at_k_vs_exact_f32"] = {"value": o.recall_at_k(exact[0], got_ids[:ns], k), "queries": ns}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_flat(oix, queries, k, n, dim, nq, self_mode=self_queries)
    index.close()
    return line


# ---- IVF ----------------------------------------------------------------------------------------------------------
def bench_ivf_set(c, n, dim, nq, k, nlist, entries, tf32_peak):
    """One shared build of the index on the device (f32 / BF16 / SQ8 forms of the same lists), then one bench entry per
    (dtype, nprobe)."""
    torch, annb200, args = c.torch, c.annb, c.args
    import gpu_setup as gs
    from annb200 import distributed as D
    from oracle import oracle as o
    kind = "correlated"
    t_setup = time.perf_counter()
    data_t = gs.correlated_gpu(n, dim, c.dev, seed=42)
    q_t = gs.subsample_with_noise_gpu(data_t, nq, seed=42)
    queries = q_t.cpu().numpy()
    dtypes = sorted({e[0] for e in entries}, key=["f32", "bf16", "sq8"].index)
    base = gs.build_ivf_parts_gpu(data_t, nlist, annb200.F32, c.local_rank, seed=42, kmeans_iters=args.kmeans_iters)
    truth_ids = None
    if c.rank == 0:
        truth_ids = gs.exact_ground_truth(data_t, q_t[:min(1000, nq)].contiguous(), k, annb200.L2, c.local_rank)
    del data_t, q_t
    torch.cuda.empty_cache()
    if c.rank == 0:
        print(f"[setup] {n}x{dim} data + IVF lists (nlist {nlist}) + exact ground truth on the GPU in {time.perf_counter() - t_setup:.1f} s", file=sys.stderr)
    out = []
    lb, le = D.list_ranges(base["offsets"], c.world)[c.rank]
    cores = os.cpu_count() or 1
    for dtype in dtypes:
        dt = {"f32": annb200.F32, "bf16": annb200.BF16, "sq8": annb200.SQ8}[dtype]
        parts = gs.requantise_parts(base, dt)
        if c.single:
            index = gs.ivf_handle_from_parts(parts, n, dim, dt, annb200.L2, c.devices)
        else:
            index = gs.ivf_handle_from_parts(parts, n, dim, dt, annb200.L2, c.local_rank, lb, le)
        apply_options(args, index)
        oi = None
        if c.rank == 0:      # host mirror of the same index contents for the oracle (parity sample + CPU baseline)
            vec = parts["vectors"]
            vec_np = (vec.view(torch.int16).cpu().numpy().view(np.uint16) if dt == annb200.BF16 else vec.cpu().numpy())
            oi = o.IvfIndex({"f32": o.F32, "bf16": o.BF16, "sq8": o.SQ8}[dtype], o.L2, n, dim, nlist, vec_np, parts["centroids"].cpu().numpy(),
                            parts["offsets"].astype(np.int64), parts["original_ids"].cpu().numpy(),
                            scales=None if parts["scales"] is None else parts["scales"].cpu().numpy())
        if dt != annb200.F32:
            del parts
            torch.cuda.empty_cache()
        for (_, nprobe) in [e for e in entries if e[0] == dtype]:
            se = Searcher(c, index, True, nq, dim, k, nprobe, queries)
            r = timed_run(c, index, se.step_device, se.step_host, nq, args.steps, args.warmup, se.flush)
            last_path = index.get_stat("last_path")
            fallback_q = index.get_stat("fallback_queries")
            dev_ids, dev_dist = se.device_result()
            scanned = scanned_local = 0
            if c.world == 1 and not c.single:          # probe statistics come with the host-buffer call of one handle
                scanned, scanned_local = index.get_stat("scanned_vectors"), index.get_stat("scanned_vectors_local")
            if c.rank != 0:
                continue
            got_ids, got_d = se.h_ids.numpy(), se.h_dist.numpy()
            esz = {"f32": 4, "bf16": 2, "sq8": 1}[dtype]
            per_vec = dim * esz
            if not scanned:                            # sharded runs: the oracle's own probe walk gives the same sum over a sample
                samp = o.ivf_search(oi, queries[:256], k, nprobe=nprobe, nthreads=cores)
                scanned = int(samp[4].sum() * (nq / 256.0))
                scanned_local = scanned // c.shards
            algo_bytes = scanned_local * per_vec          # sum over queries of the probed lists' bytes (this rank's lists)
            hbm = {"algorithmic_bytes_per_launch": int(algo_bytes), "algorithmic_gbs": algo_bytes / r["dom_s"] / 1e9 if r["dom_s"] else None,
                   "peak_gbs": c.peaks["hbm_gbs"],
                   "note": "a list tile loaded once serves up to 128 queries of the batch, so the per-query (algorithmic) byte rate exceeds the HBM peak by design; "
                           "dram_gbs_from_traffic is what the kernel really moved (ncu), when a capture of this workload is committed"}
            if last_path == 2:
                try:
                    ivf_kind = index.get_stat("tc_kind")
                except Exception:
                    ivf_kind = 0
                roofline = tensor_roofline(c, dtype, 2.0 * scanned_local * dim, r["dom_s"], "ivf grouped list scan + top-k' select (ivf_tc_kernel)", tf32_peak,
                                           bf16_terms=3, f32_kind=ivf_kind)
                roofline["algorithmic_flops_per_launch"] = 2.0 * scanned_local * dim
                roofline["executed"]["note"] = "padded (list x 128-query group) tiles execute more MMA work than the algorithmic count"
                tiles = index.get_stat("tc_scan_tiles") if (c.world == 1 and not c.single) else 0
                if tiles and r["dom_s"]:
                    # what the tensor pipe really executed: every task multiplies whole 128-row tiles of its list by a 128-row query group
                    # (groups hold fewer than 128 queries at this batch size), K padded to the 128-byte slab, all MMA terms
                    slab = {"f32": 64 if ivf_kind == 3 else 32, "bf16": 64, "sq8": 128}[dtype]
                    kp = -(-dim // slab) * slab
                    ex = roofline["executed"]
                    padded = tiles * 128.0 * 128.0 * kp * 2.0 * ex["mma_terms_per_element"]
                    ex["padded"] = {"tiles_per_launch": int(tiles), "tflops": padded / r["dom_s"] / 1e12,
                                    "frac_of_pipe_peak": padded / r["dom_s"] / 1e12 / ex["pipe_peak"],
                                    "query_group_fill": (2.0 * scanned_local * dim * ex["mma_terms_per_element"]) / padded,
                                    "note": "tiles counted by the library (stat tc_scan_tiles) for the same batch; the padded figure is the tensor-pipe utilisation, "
                                            "the algorithmic one (frac) is what the metric pays for"}
                roofline["hbm_view"] = hbm
                roofline["path"] = "tensor (tcgen05)"
            else:
                achieved = algo_bytes / r["dom_s"] / 1e9
                roofline = {"bound": "hbm", "achieved": achieved, "peak": c.peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / c.peaks["hbm_gbs"], "traffic": None,
                            "kernel": "ivf list scan (CUDA cores)", "kernel_ms": r["dom_s"] * 1e3, "peak_source": f"{c.peak_src} copy bandwidth",
                            "algorithmic_bytes_per_launch": int(algo_bytes), "path": "simt (CUDA cores)"}
            key = f"ivf {dtype} n={n} dim={dim} nq={nq} nlist={nlist} nprobe={nprobe}"
            if key in c.traffic and c.shards == 1:
                roofline["traffic"] = c.traffic[key]["dram_bytes"]
                roofline["traffic_source"] = c.traffic[key]["source"]
                if "hbm_view" in roofline and r["dom_s"]:
                    roofline["hbm_view"]["dram_gbs_from_traffic"] = c.traffic[key]["dram_bytes"] / r["dom_s"] / 1e9
                    roofline["hbm_view"]["frac_of_hbm_peak"] = c.traffic[key]["dram_bytes"] / r["dom_s"] / 1e9 / c.peaks["hbm_gbs"]
            entry = {"metric": f"QPS ivf {dtype} euclidean nlist={nlist} nprobe={nprobe} k={k}", "value": r["qps"], "unit": "queries/s", "n_gpus": c.shards,
                     "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                     "dtype": {"f32": "f32 (3xFP16 / 3xTF32 split-precision select + f32 exact re-rank)", "bf16": "bf16 (f32 accumulate)", "sq8": "int8 (i32 accumulate)"}[dtype],
                     "config": {"workload": f"IVF {dtype} euclidean, {n}x{dim} {kind} synthetic, nlist={nlist}, nprobe={nprobe}, {nq}-query batch, k={k}",
                                "sharding": (f"inverted lists over {c.shards} GPUs, " +
                                             ("one process, peer copies" if c.single else "one process per GPU, probe exchange + one NCCL all-gather; device loop: verdict of step i read after step i+1 is enqueued"))
                                if c.shards > 1 else "none",
                                "algorithmic_bytes_per_query": scanned * per_vec / nq, "l2_policy": "inputs larger than L2"},
                     "clocks": r["clocks"],
                     "e2e": {"value": r["e2e_qps"], "unit": "queries/s", "h2d_bytes_per_step": int(nq * dim * 4),
                             "d2h_bytes_per_step": int(nq * k * 12 + (nq * 4 if c.world == 1 else 0)), "ms_per_step": r["e2e_s"] * 1e3},
                     "gpu_launches": r["launches"] + se.extra_launches, "roofline": roofline, "fallback_queries_total": int(fallback_q)}
            if se.pair is not None:
                entry["shard_refined_queries_total"] = int(sum(s.refined_queries for s in se.pair))
            if c.shards > 1:
                entry["device_vs_e2e"] = {"ids_equal": bool(np.array_equal(dev_ids, got_ids)),
                                          "dist_bits_equal": bool(np.array_equal(dev_dist.view(np.uint32), got_d.view(np.uint32)))}
            ns = min(64, nq)
            ref = o.ivf_search(oi, queries[:ns], k, nprobe=nprobe, nthreads=cores)
            entry["parity_sample"] = {"queries": ns, "ids_equal": bool(np.array_equal(got_ids[:ns], ref[0])),
                                      "dist_bits_equal": bool(np.array_equal(got_d[:ns].view(np.uint32), ref[1].view(np.uint32))),
                                      "against": "CPU oracle (oracle/oracle.c) on the same index contents" + (f", merged over {c.shards} shards" if c.shards > 1 else "")}
            if dtype == "sq8":   # integer distances tie: ids are compared per tie class (see tie_classes_equal)
                entry["parity_sample"]["tie_classes_equal"] = bool(tie_classes_equal(np.asarray(got_ids[:ns]).astype(np.int64), got_d[:ns], np.asarray(ref[0]).astype(np.int64), ref[1]))
            entry["recall_at_k_vs_exact_f32"] = {"value": o.recall_at_k(truth_ids, got_ids[:truth_ids.shape[0]], k), "queries": int(truth_ids.shape[0])}
            if not args.no_cpu_baseline:
                per_query_s = (nprobe * n / nlist + nlist) * dim / 3.0e9
                nsb = int(max(cores, min(nq, 8.0 / max(per_query_s, 1e-9))))
                o.ivf_search(oi, queries[:max(8, min(nsb, cores))], k, nprobe=nprobe, nthreads=cores)
                t0 = time.perf_counter()
                o.ivf_search(oi, queries[:nsb], k, nprobe=nprobe, nthreads=cores)
                dtc = time.perf_counter() - t0
                entry["cpu_baseline"] = {"value": nsb / dtc, "unit": "queries/s", "cores": cores, "kind": "port",
                                         "sample": f"{nsb} of {nq} queries against the same index, {dtc:.2f} s wall"}
            out.append(entry)
        index.close()
        del oi
    return out


def run_b200(args):
    c = init_ctx(args)
    tf32_peak = measure_tf32_peak(c) if c.rank == 0 else None
    line = None
    wl = args.workload
    nq = args.nq
    if wl in ("all", "flat"):
        line = bench_flat(c, args.n or 1_000_000, args.dim or 128, nq, args.k or 10, args.metric or "cosine", args.dtype, args.self_queries, tf32_peak)
    if wl in ("all", "ivf"):
        if wl == "all":
            entries = [(e.split(":")[0], int(e.split(":")[1])) for e in args.ivf_set.split(",")]
            ivf = bench_ivf_set(c, 10_000_000, 128, nq, 10, 4096, entries, tf32_peak)
        else:
            ivf = bench_ivf_set(c, args.n or 10_000_000, args.dim or 128, nq, args.k or 10, args.nlist, [(args.dtype, args.nprobe)], tf32_peak)
        if c.rank == 0:
            if wl == "ivf":
                line = dict(ivf[0])
                line.update({"higher_is_better": True, "scaling": "strong", "vs_baseline": None, "data": "synthetic"})
            else:
                line["ivf"] = ivf
    if wl in ("all", "c5"):
        if wl == "all":
            c5 = bench_flat(c, 2_000_000, 50, nq, 15, "euclidean", "f32", True, tf32_peak, on_gpu_data=True)
        else:
            c5 = bench_flat(c, args.n or 2_000_000, args.dim or 50, nq, args.k or 15, "euclidean", args.dtype, True, tf32_peak, on_gpu_data=True)
        if c.rank == 0:
            if wl == "c5":
                line = c5
            else:
                line["c5"] = c5
    if c.rank == 0:
        c.sampler.stop()
        line["tf32_gemm_tflops_measured_in_run"] = tf32_peak
        print(json.dumps(line))
    if c.world > 1:
        c.dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
