timeout 1100 python -m pytest tests -x -q -m gpu > gpurun_out/V4_pytest.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/V4_pytest.log
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" > gpurun_out/V4_$name.json 2> gpurun_out/V4_$name.err; python tools/show_bench.py gpurun_out/V4_$name.json 2>&1 | cut -c1-250; }
run ivf_gm1 --workload ivf --ivf-set f32:32
run ivf_gm0 --workload ivf --ivf-set f32:32 --option ivf_coarse_gm=0
IVF="python bench.py --workload ivf --ivf-set f32:32 --steps 2 --warmup 2 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/V4_launches_ivf_gm1.csv $IVF > gpurun_out/V4_ncu1.log 2>&1
python profiles/launch_summary.py gpurun_out/V4_launches_ivf_gm1.csv 2>&1 | head -14
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:coarse_select_gm -c 1 -o gpurun_out/V4_coarse -f $IVF > gpurun_out/V4_ncu2.log 2>&1
python profiles/ncu_top.py gpurun_out/V4_coarse.ncu-rep 40 > gpurun_out/V4_coarse.txt 2>&1
timeout 300 python tools/shard_emulate.py --workload ivf --world 8 2>&1 | tail -9 | head -5
