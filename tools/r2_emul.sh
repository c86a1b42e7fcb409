set -x
timeout 300 python tools/shard_emulate.py --workload ivf --world 8 --nprobe 32 > gpurun_out/E_ivf8.log 2>&1; tail -12 gpurun_out/E_ivf8.log
timeout 300 python tools/shard_emulate.py --workload flat --world 8 > gpurun_out/E_flat8.log 2>&1; tail -12 gpurun_out/E_flat8.log
timeout 300 python tools/shard_emulate.py --workload c5 --world 8 > gpurun_out/E_c58.log 2>&1; tail -12 gpurun_out/E_c58.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/E_ivf8_launches.csv python tools/shard_emulate.py --workload ivf --world 8 --nprobe 32 --steps 2 --warmup 2 > gpurun_out/E_ncu_ivf8.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/E_flat8_launches.csv python tools/shard_emulate.py --workload flat --world 8 --steps 2 --warmup 2 > gpurun_out/E_ncu_flat8.log 2>&1
