# Validation of HEAD as the driver will run it: GPU tests, smoke, bench (driver flags), reference arm; wall time of each.
t0=$(date +%s)
timeout 1100 python -m pytest tests -x -q -m gpu > gpurun_out/V1_pytest.log 2>&1; echo "pytest rc $? $(( $(date +%s)-t0 )) s"; tail -3 gpurun_out/V1_pytest.log
t0=$(date +%s)
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "smoke $(( $(date +%s)-t0 )) s"
t0=$(date +%s)
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/V1_bench.json 2> gpurun_out/V1_bench.err; echo "bench rc $? $(( $(date +%s)-t0 )) s"
python tools/show_bench.py gpurun_out/V1_bench.json 2>&1 | cut -c1-250
t0=$(date +%s)
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/V1_ref.json 2> gpurun_out/V1_ref.err; echo "ref rc $? $(( $(date +%s)-t0 )) s"; cut -c1-300 gpurun_out/V1_ref.json
for w in ivf flat c5; do timeout 300 python tools/shard_emulate.py --workload $w --world 8 > gpurun_out/V1_emul_$w.log 2>&1; tail -14 gpurun_out/V1_emul_$w.log; done
