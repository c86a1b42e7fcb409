timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/H8_bench.json 2> gpurun_out/H8_bench.err; echo "bench rc $?"; tail -3 gpurun_out/H8_bench.err
python tools/show_bench.py gpurun_out/H8_bench.json | cut -c1-260
