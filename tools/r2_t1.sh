timeout 900 python -m pytest tests/test_gpu_tensor.py -x -q -m gpu > gpurun_out/T1_pytest.log 2>&1; echo "pytest rc $?"; tail -30 gpurun_out/T1_pytest.log
timeout 800 python tools/k_dim_sweep.py > gpurun_out/K2_sweep.txt 2>&1; cat gpurun_out/K2_sweep.txt
