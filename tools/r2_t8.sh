timeout 900 python -m pytest tests/test_gpu_flat.py tests/test_gpu_ivf.py tests/test_gpu_tensor.py -x -q -m gpu > gpurun_out/T8_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/T8_pytest.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/T8_ivf_launches.csv python bench.py --workload ivf --ivf-set f32:32 --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/T8_ncu.log 2>&1
python profiles/launch_summary.py gpurun_out/T8_ivf_launches.csv | head -9
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/T8_flat_launches.csv python bench.py --workload flat --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/T8_ncu2.log 2>&1
python profiles/launch_summary.py gpurun_out/T8_flat_launches.csv | head -6
