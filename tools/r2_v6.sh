timeout 1100 python -m pytest tests -x -q -m gpu > gpurun_out/V6_pytest.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/V6_pytest.log
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" > gpurun_out/V6_$name.json 2> gpurun_out/V6_$name.err; python tools/show_bench.py gpurun_out/V6_$name.json 2>&1 | cut -c1-250; }
run ivf --workload ivf --ivf-set f32:32
IVF="python bench.py --workload ivf --ivf-set f32:32 --steps 2 --warmup 2 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/V6_launches_ivf.csv $IVF > gpurun_out/V6_ncu1.log 2>&1
python profiles/launch_summary.py gpurun_out/V6_launches_ivf.csv 2>&1 | head -14
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"flat_tc_kernel" -c 1 -o gpurun_out/V6_dense -f $IVF > gpurun_out/V6_ncu2.log 2>&1
python profiles/ncu_top.py gpurun_out/V6_dense.ncu-rep 40 > gpurun_out/V6_dense.txt 2>&1
for w in ivf flat; do timeout 300 python tools/shard_emulate.py --workload $w --world 8 2>&1 | tail -10 | head -6; done
