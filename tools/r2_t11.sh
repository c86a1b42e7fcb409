timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/T11_pytest.log 2>&1; echo "pytest rc $?"; tail -25 gpurun_out/T11_pytest.log
timeout 500 python bench.py > gpurun_out/T11_bench.json 2> gpurun_out/T11_bench.err; echo "bench rc $?"; tail -2 gpurun_out/T11_bench.err
python tools/show_bench.py gpurun_out/T11_bench.json 2>&1 | cut -c1-250
