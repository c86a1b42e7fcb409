timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_flat.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/T15_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/T15_pytest.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/T15_flat_launches.csv python bench.py --workload flat --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/T15_ncu2.log 2>&1
python profiles/launch_summary.py gpurun_out/T15_flat_launches.csv 2>/dev/null | head -6
