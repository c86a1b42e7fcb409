timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/T2_pytest.log 2>&1; echo "pytest rc $?"; tail -15 gpurun_out/T2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --ivf-set f32:32 > gpurun_out/T2_bench_2gpu.json 2> gpurun_out/T2_bench_2gpu.err; echo "bench rc $?"; tail -3 gpurun_out/T2_bench_2gpu.err
python tools/show_bench.py gpurun_out/T2_bench_2gpu.json | cut -c1-260
