timeout 1100 python -m pytest tests -x -q -m gpu > gpurun_out/F1_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/F1_pytest.log
timeout 500 python bench.py > gpurun_out/F1_bench.json 2> gpurun_out/F1_bench.err; echo "bench rc $?"; tail -2 gpurun_out/F1_bench.err
python tools/show_bench.py gpurun_out/F1_bench.json 2>&1 | cut -c1-250
for dt in bf16 sq8; do timeout 300 python bench.py --workload flat --dtype $dt > gpurun_out/F1_flat_$dt.json 2> gpurun_out/F1_flat_$dt.err; python tools/show_bench.py gpurun_out/F1_flat_$dt.json 2>&1 | cut -c1-250; done
