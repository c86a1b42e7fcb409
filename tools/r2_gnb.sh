N=$1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/H${N}_bench.json 2> gpurun_out/H${N}_bench.err; echo "bench rc $?"; tail -2 gpurun_out/H${N}_bench.err
python tools/show_bench.py gpurun_out/H${N}_bench.json | cut -c1-260
