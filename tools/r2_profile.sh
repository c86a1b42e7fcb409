# Round-2 ncu evidence (one GPU): launch lists of the timed regions and full captures of the three dominant kernels.
# Every ncu command runs right behind a plain run of the same command line (B200_PROFILING.md).
set -x
TM="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__ops_path_tensor_src_tf32_dst_fp32.sum,sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32.sum"
IVF="python bench.py --workload ivf --dtype f32 --nprobe 32 --steps 2 --warmup 2 --no-cpu-baseline"
FLAT="python bench.py --workload flat --steps 2 --warmup 2 --no-cpu-baseline"
STREAM="python bench.py --workload ivf --dtype f32 --nprobe 32 --nq 1000 --list-major 0 --steps 2 --warmup 2 --no-cpu-baseline"
$IVF > gpurun_out/P_ivf_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/P_launches_ivf_f32.csv $IVF > gpurun_out/P_ncu1.log 2>&1
$FLAT > gpurun_out/P_flat_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/P_launches_flat_f32.csv $FLAT > gpurun_out/P_ncu2.log 2>&1
ncu --profile-from-start off --set full --metrics $TM --clock-control none --import-source on -k regex:ivf_tc_kernel -c 1 -o gpurun_out/P_ivf_tc_f32 -f $IVF > gpurun_out/P_ncu3.log 2>&1
ncu --profile-from-start off --set full --metrics $TM --clock-control none --import-source on -k regex:flat_tc_kernel -c 1 -o gpurun_out/P_flat_tc_f32 -f $FLAT > gpurun_out/P_ncu4.log 2>&1
$STREAM > gpurun_out/P_stream_plain.log 2>&1 && ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ivf_stream_kernel -c 1 -o gpurun_out/P_ivf_stream_f32 -f $STREAM > gpurun_out/P_ncu5.log 2>&1
ls -la gpurun_out/P_*
