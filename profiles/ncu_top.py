#!/usr/bin/env python
"""Summarise an .ncu-rep: headline raw metrics + the instructions that collect the most stall samples.
usage: python profiles/ncu_top.py gpurun_out/prof.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__cycles_elapsed.max", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32.sum", "sm__ops_path_tensor_src_tf32_dst_fp32.sum", "sm__ops_path_tensor_src_int8.sum",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum", "sm__ops_path_tensor_src_fp16_dst_fp32.sum", "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second"]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"  {h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {k: sum(f(r, k) for r in data) for k in stalls}
print("stall mix:", ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print(f"top {N} instructions by samples (of {int(tot)}):")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:N]:
    s = sorted(((k, f(r, k)) for k in stalls), key=lambda kv: -kv[1])[0]
    print(f"  {r[idx['Address']][-5:]} {100*f(r,'# Samples')/tot:6.2f}% exec={int(f(r,'Instructions Executed')):>10} {r[idx['Source']][:64]:<64} {s[0][6:]}")
