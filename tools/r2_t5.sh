timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/T5_pytest.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/T5_pytest.log
timeout 300 python tools/shard_emulate.py --workload ivf --world 8 --nprobe 32 2>&1 | tail -9 | head -5
timeout 300 python tools/shard_emulate.py --workload flat --world 8 2>&1 | tail -9 | head -5
