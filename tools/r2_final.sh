# Round-2 final evidence, one GPU.  Every ncu command runs right behind a plain run of the same command line.
set -x
timeout 1100 python -m pytest tests -x -q -m gpu > gpurun_out/Z_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/Z_pytest.log
timeout 500 python bench.py > gpurun_out/Z_bench_1gpu.json 2> gpurun_out/Z_bench_1gpu.err; echo "bench rc $?"
python tools/show_bench.py gpurun_out/Z_bench_1gpu.json 2>&1 | cut -c1-250
for dt in bf16 sq8; do timeout 300 python bench.py --workload flat --dtype $dt > gpurun_out/Z_flat_$dt.json 2> /dev/null; python tools/show_bench.py gpurun_out/Z_flat_$dt.json 2>&1 | cut -c1-250; done
timeout 300 python bench.py --workload flat --option tc_f32_fp16=0 --no-cpu-baseline > gpurun_out/Z_flat_tf32.json 2> /dev/null; python tools/show_bench.py gpurun_out/Z_flat_tf32.json 2>&1 | cut -c1-250
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/Z_reference_arm.json 2> /dev/null; cut -c1-300 gpurun_out/Z_reference_arm.json
timeout 800 python tools/k_dim_sweep.py > gpurun_out/Z_k_dim_sweep.txt 2>&1; cat gpurun_out/Z_k_dim_sweep.txt
for w in ivf flat c5; do timeout 300 python tools/shard_emulate.py --workload $w --world 8 > gpurun_out/Z_emul_$w.log 2>&1; tail -12 gpurun_out/Z_emul_$w.log; done
TM="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__ops_path_tensor_src_tf32_dst_fp32.sum,sm__ops_path_tensor_src_fp16_dst_fp32.sum,sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32.sum"
IVF="python bench.py --workload ivf --ivf-set f32:32 --steps 2 --warmup 2 --no-cpu-baseline"
FLAT="python bench.py --workload flat --steps 2 --warmup 2 --no-cpu-baseline"
STREAM="python bench.py --workload ivf --dtype f32 --nprobe 32 --nq 1000 --list-major 0 --steps 2 --warmup 2 --no-cpu-baseline"
$STREAM > gpurun_out/Z_stream_plain.json 2> /dev/null; python tools/show_bench.py gpurun_out/Z_stream_plain.json 2>&1 | cut -c1-250
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/Z_launches_ivf_f32.csv $IVF > gpurun_out/Z_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/Z_launches_flat_f32.csv $FLAT > gpurun_out/Z_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/Z_launches_emul_ivf8.csv python tools/shard_emulate.py --workload ivf --world 8 --steps 2 --warmup 2 > gpurun_out/Z_ncu3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/Z_launches_emul_flat8.csv python tools/shard_emulate.py --workload flat --world 8 --steps 2 --warmup 2 > gpurun_out/Z_ncu4.log 2>&1
ncu --profile-from-start off --set full --metrics $TM --clock-control none --import-source on -k regex:ivf_tc_kernel -c 1 -o gpurun_out/Z_ivf_tc_f32 -f $IVF > gpurun_out/Z_ncu5.log 2>&1
ncu --profile-from-start off --set full --metrics $TM --clock-control none --import-source on -k regex:flat_tc_kernel -c 1 -o gpurun_out/Z_flat_tc_f32 -f $FLAT > gpurun_out/Z_ncu6.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ivf_stream_kernel -c 1 -o gpurun_out/Z_ivf_stream_f32 -f $STREAM > gpurun_out/Z_ncu7.log 2>&1
for k in ivf_tc_f32 flat_tc_f32 ivf_stream_f32; do python profiles/ncu_top.py gpurun_out/Z_$k.ncu-rep 30 > gpurun_out/Z_$k.txt 2>&1; done
ls -la gpurun_out/Z_*
