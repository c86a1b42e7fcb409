timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_ivf.py tests/test_gpu_flat.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/T3_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/T3_pytest.log
timeout 300 python bench.py --workload ivf --ivf-set f32:32,bf16:32 --no-cpu-baseline > gpurun_out/T3_ivf.json 2>/dev/null; python tools/show_bench.py gpurun_out/T3_ivf.json | cut -c1-200
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/T3_ivf_launches.csv python bench.py --workload ivf --ivf-set f32:32 --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/T3_ncu.log 2>&1
python profiles/launch_summary.py gpurun_out/T3_ivf_launches.csv | head -12
