timeout 1100 python -m pytest tests -x -q -m gpu > gpurun_out/V8_pytest.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/V8_pytest.log
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" > gpurun_out/V8_$name.json 2> gpurun_out/V8_$name.err; python tools/show_bench.py gpurun_out/V8_$name.json 2>&1 | cut -c1-250; }
run ivf_walk1 --workload ivf --ivf-set f32:8,f32:32,f32:128,bf16:32,sq8:32
run ivf_walk0 --workload ivf --ivf-set f32:32 --option ivf_coarse_walk=0
IVF="python bench.py --workload ivf --ivf-set f32:32 --steps 2 --warmup 2 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/V8_launches_ivf.csv $IVF > gpurun_out/V8_ncu1.log 2>&1
python profiles/launch_summary.py gpurun_out/V8_launches_ivf.csv 2>&1 | head -14
timeout 300 python tools/shard_emulate.py --workload ivf --world 8 2>&1 | tail -10 | head -6
