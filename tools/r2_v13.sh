run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > gpurun_out/V13_$name.json 2> gpurun_out/V13_$name.err; python tools/show_bench.py gpurun_out/V13_$name.json 2>&1 | cut -c1-250; }
run flat_w --workload flat
ANNB200_LIB=ann-search-rs_b200/lib/libannb200_alt.so run flat_v --workload flat
run flat_w2 --workload flat
ANNB200_LIB=ann-search-rs_b200/lib/libannb200_alt.so run flat_v2 --workload flat
timeout 600 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_flat.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/shard_emulate.py --workload flat --world 8 2>&1 | tail -10 | head -3
ANNB200_LIB=ann-search-rs_b200/lib/libannb200_alt.so timeout 300 python tools/shard_emulate.py --workload flat --world 8 2>&1 | tail -10 | head -3
