ANNB200_LIB=ann-search-rs_b200/lib/libannb200_counters.so timeout 300 python tools/ivf_cycles.py 10000000 10000 32 f32 2>&1 | tail -1
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" > gpurun_out/V10_$name.json 2> gpurun_out/V10_$name.err; python tools/show_bench.py gpurun_out/V10_$name.json 2>&1 | cut -c1-250; }
for dt in f32 bf16 sq8; do
run ivf_${dt}_pf1 --workload ivf --dtype $dt --nprobe 32
run ivf_${dt}_pf0 --workload ivf --dtype $dt --nprobe 32 --option ivf_task_prefetch=0
done
run ivf_f32_np8 --workload ivf --dtype f32 --nprobe 8
run ivf_f32_np128 --workload ivf --dtype f32 --nprobe 128
timeout 600 python -m pytest tests/test_gpu_ivf.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/shard_emulate.py --workload ivf --world 8 2>&1 | tail -10 | head -6
