set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
IVF="python bench.py --workload ivf --dtype f32 --nprobe 32 --steps 2 --warmup 2 --no-cpu-baseline"
$IVF > gpurun_out/Q_ivf_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/Q_launches_ivf_f32.csv $IVF > gpurun_out/Q_ncu1.log 2>&1
python profiles/launch_summary.py gpurun_out/Q_launches_ivf_f32.csv
for nq in 1000 256 1; do ANNB_AB_FIXED=ivf_list_major=0 python tools/ivf_ab.py 10000000 $nq 32 f32,bf16,sq8 ivf_stream 0,1 2>&1 | grep "rep 1"; done
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_all_2.json 2> gpurun_out/r2_bench_all_2.err; echo bench rc=$?
