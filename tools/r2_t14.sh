ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:rerank_kernel -c 1 -o gpurun_out/R2_rerank_flat -f python bench.py --workload flat --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/R2_ncu.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:rerank_kernel -c 1 -o gpurun_out/R2_rerank_ivf -f python bench.py --workload ivf --ivf-set f32:32 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/R2_ncu2.log 2>&1
ls -la gpurun_out/R2_*
