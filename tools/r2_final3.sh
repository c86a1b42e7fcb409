# Round-2 closing evidence, one GPU.  Every ncu command runs right behind a plain run of the same command line.
set -x
timeout 1100 python -m pytest tests -x -q -m gpu > gpurun_out/F_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/F_pytest.log
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/F_bench_1gpu.json 2> gpurun_out/F_bench_1gpu.err; echo "bench rc $?"
python tools/show_bench.py gpurun_out/F_bench_1gpu.json 2>&1 | cut -c1-250
for dt in bf16 sq8; do timeout 300 python bench.py --workload flat --dtype $dt > gpurun_out/F_flat_$dt.json 2> /dev/null; python tools/show_bench.py gpurun_out/F_flat_$dt.json 2>&1 | cut -c1-250; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/F_reference_arm.json 2> /dev/null; cut -c1-300 gpurun_out/F_reference_arm.json
for w in ivf flat c5; do timeout 300 python tools/shard_emulate.py --workload $w --world 8 > gpurun_out/F_emul_$w.log 2>&1; tail -13 gpurun_out/F_emul_$w.log; done
TM="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__ops_path_tensor_src_tf32_dst_fp32.sum,sm__ops_path_tensor_src_fp16_dst_fp32.sum,sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32.sum"
IVF="python bench.py --workload ivf --ivf-set f32:32 --steps 2 --warmup 2 --no-cpu-baseline"
FLAT="python bench.py --workload flat --steps 2 --warmup 2 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/F_launches_ivf_f32.csv $IVF > gpurun_out/F_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/F_launches_flat_f32.csv $FLAT > gpurun_out/F_ncu2.log 2>&1
python profiles/launch_summary.py gpurun_out/F_launches_ivf_f32.csv > gpurun_out/F_launches_ivf_f32.txt; cat gpurun_out/F_launches_ivf_f32.txt
python profiles/launch_summary.py gpurun_out/F_launches_flat_f32.csv > gpurun_out/F_launches_flat_f32.txt; cat gpurun_out/F_launches_flat_f32.txt
ncu --profile-from-start off --set full --metrics $TM --clock-control none --import-source on -k regex:ivf_tc_kernel -c 1 -o gpurun_out/F_ivf_tc_f32 -f $IVF > gpurun_out/F_ncu5.log 2>&1
ncu --profile-from-start off --set full --metrics $TM --clock-control none --import-source on -k regex:coarse_select_gm -c 1 -o gpurun_out/F_coarse_select_gm -f $IVF > gpurun_out/F_ncu6.log 2>&1
ncu --profile-from-start off --set full --metrics $TM --clock-control none --import-source on -k regex:flat_tc_kernel -c 1 -o gpurun_out/F_dense_f16 -f $IVF > gpurun_out/F_ncu7.log 2>&1
ncu --profile-from-start off --set full --metrics $TM --clock-control none --import-source on -k regex:flat_tc_kernel -c 1 -o gpurun_out/F_flat_tc_f32 -f $FLAT > gpurun_out/F_ncu8.log 2>&1
for k in ivf_tc_f32 coarse_select_gm dense_f16 flat_tc_f32; do python profiles/ncu_top.py gpurun_out/F_$k.ncu-rep 30 > gpurun_out/F_$k.txt 2>&1; rm -f gpurun_out/F_$k.ncu-rep; done
ls -la gpurun_out/F_*
