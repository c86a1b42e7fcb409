set -x
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:rerank_kernel -c 1 -o gpurun_out/R_rerank_ivf -f python bench.py --workload ivf --dtype f32 --nprobe 32 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/R_ncu1.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:rerank_kernel -c 1 -o gpurun_out/R_rerank_flat -f python bench.py --workload flat --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/R_ncu2.log 2>&1
python profiles/ncu_top.py gpurun_out/R_rerank_ivf.ncu-rep 40 > gpurun_out/R_rerank_ivf.txt 2>&1
python profiles/ncu_top.py gpurun_out/R_rerank_flat.ncu-rep 40 > gpurun_out/R_rerank_flat.txt 2>&1
ls -la gpurun_out/R_*
