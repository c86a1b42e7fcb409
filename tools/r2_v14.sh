run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > gpurun_out/V14_$name.json 2> gpurun_out/V14_$name.err; python tools/show_bench.py gpurun_out/V14_$name.json 2>&1 | cut -c1-250; }
run flat_unit --workload flat
ANNB200_LIB=ann-search-rs_b200/lib/libannb200_alt.so run flat_prev --workload flat
run flat_unit2 --workload flat
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/shard_emulate.py --workload flat --world 8 2>&1 | tail -10 | head -3
