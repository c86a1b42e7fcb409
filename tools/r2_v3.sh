IVF="python bench.py --workload ivf --ivf-set f32:32 --steps 2 --warmup 2 --no-cpu-baseline"
for k in coarse_select_gm rerank_kernel; do
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/V3_$k -f $IVF > gpurun_out/V3_ncu_$k.log 2>&1
tail -2 gpurun_out/V3_ncu_$k.log
python profiles/ncu_top.py gpurun_out/V3_$k.ncu-rep 70 > gpurun_out/V3_$k.txt 2>&1
done
ls -la gpurun_out/V3_*
