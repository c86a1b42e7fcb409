N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/G${N}_bench.json 2> gpurun_out/G${N}_bench.err; echo "bench rc $?"; tail -2 gpurun_out/G${N}_bench.err
python tools/show_bench.py gpurun_out/G${N}_bench.json | cut -c1-260
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>/dev/null | cut -c1-400
