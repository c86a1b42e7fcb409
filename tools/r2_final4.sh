# Round-2 closing evidence after the epilogue rework (unit-row cosine operand, uniform-scale L2 operands), one GPU.
set -x
timeout 1100 python -m pytest tests -q -m gpu > gpurun_out/F4_pytest.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/F4_pytest.log
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/F4_bench_1gpu.json 2> gpurun_out/F4_bench_1gpu.err; echo "bench rc $?"
python tools/show_bench.py gpurun_out/F4_bench_1gpu.json 2>&1 | cut -c1-250
for w in ivf flat c5; do timeout 300 python tools/shard_emulate.py --workload $w --world 8 > gpurun_out/F4_emul_$w.log 2>&1; tail -10 gpurun_out/F4_emul_$w.log | head -6; done
TM="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__ops_path_tensor_src_fp16_dst_fp32.sum"
IVF="python bench.py --workload ivf --ivf-set f32:32 --steps 2 --warmup 2 --no-cpu-baseline"
FLAT="python bench.py --workload flat --steps 2 --warmup 2 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/F4_launches_ivf_f32.csv $IVF > gpurun_out/F4_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/F4_launches_flat_f32.csv $FLAT > gpurun_out/F4_ncu2.log 2>&1
python profiles/launch_summary.py gpurun_out/F4_launches_ivf_f32.csv > gpurun_out/F4_launches_ivf_f32.txt; cat gpurun_out/F4_launches_ivf_f32.txt
python profiles/launch_summary.py gpurun_out/F4_launches_flat_f32.csv > gpurun_out/F4_launches_flat_f32.txt; cat gpurun_out/F4_launches_flat_f32.txt
ncu --profile-from-start off --set full --metrics $TM --clock-control none --import-source on -k regex:ivf_tc_kernel -c 1 -o gpurun_out/F4_ivf_tc_f32 -f $IVF > gpurun_out/F4_ncu5.log 2>&1
ncu --profile-from-start off --set full --metrics $TM --clock-control none --import-source on -k regex:flat_tc_kernel -c 1 -o gpurun_out/F4_flat_tc_f32 -f $FLAT > gpurun_out/F4_ncu8.log 2>&1
for k in ivf_tc_f32 flat_tc_f32; do python profiles/ncu_top.py gpurun_out/F4_$k.ncu-rep 30 > gpurun_out/F4_$k.txt 2>&1; rm -f gpurun_out/F4_$k.ncu-rep; done
