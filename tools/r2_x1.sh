run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" > gpurun_out/X1_$name.json 2> gpurun_out/X1_$name.err; python - <<P
import json
try:
    d=json.load(open("gpurun_out/X1_$name.json")); r=d.get("roofline",{})
    print("$name", "ms", round(d["ms_per_step"],3), "kern_ms", round(r.get("kernel_ms",0),3), "frac", round(r.get("frac",0),3), "fallback", d.get("fallback_queries_total"), "parity", d.get("parity_sample",{}).get("ids_equal"))
except Exception as e: print("$name", "failed", e)
P
}
run c5_base --workload c5
run c5_kp16 --workload c5 --tc-candidates 16
run c5_lo --workload c5 --option tc_f32_lo_smem=1
run c5_kp16_lo --workload c5 --tc-candidates 16 --option tc_f32_lo_smem=1
run flat_base --workload flat
run flat_lo --workload flat --option tc_f32_lo_smem=1
run flat_strided --workload flat --option tc_strided=1
timeout 300 python tools/shard_emulate.py --workload c5 --world 8 --option tc_f32_lo_smem=1 2>&1 | tail -9 | head -5
timeout 300 python tools/shard_emulate.py --workload c5 --world 8 --option tc_candidates=16 2>&1 | tail -9 | head -5
