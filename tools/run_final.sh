# Round-end measurement set (one GPU): GPU tests, the bench lines of profiles/, ncu launch lists and one full capture.
set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/F_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/F_pytest.log
timeout 300 python bench.py > gpurun_out/F_flat_f32.json 2> gpurun_out/F_flat_f32.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/F_flat_f32_reference.json 2> gpurun_out/F_flat_f32_reference.err
timeout 200 python bench.py --dtype bf16 --no-cpu-baseline > gpurun_out/F_flat_bf16.json 2> gpurun_out/F_flat_bf16.err
timeout 200 python bench.py --dtype sq8 --no-cpu-baseline > gpurun_out/F_flat_sq8.json 2> gpurun_out/F_flat_sq8.err
timeout 300 python bench.py --n 2000000 --dim 50 --k 15 --metric euclidean --self-queries > gpurun_out/F_c5_self.json 2> gpurun_out/F_c5_self.err
for np in 8 16 32 64 128; do timeout 300 python bench.py --workload ivf --nprobe $np --recall --no-cpu-baseline > gpurun_out/F_ivf_f32_np$np.json 2> gpurun_out/F_ivf_f32_np$np.err; done
timeout 400 python bench.py --workload ivf --nprobe 32 --recall > gpurun_out/F_ivf_f32_np32_cpu.json 2> gpurun_out/F_ivf_f32_np32_cpu.err
for dt in bf16 sq8; do timeout 300 python bench.py --workload ivf --dtype $dt --recall --no-cpu-baseline > gpurun_out/F_ivf_$dt.json 2> gpurun_out/F_ivf_$dt.err; done
for f in gpurun_out/F_*.json; do python -c "
import json,sys
try:
    d=json.load(open('$f'))
    r=d.get('roofline',{})
    print('$f'.split('/')[-1], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'kern_ms', round(r.get('kernel_ms',0),3), 'frac', round(r.get('frac',0),3), d.get('recall_at_k_vs_exact_f32',{}).get('value'), d.get('parity_sample'), 'uncert', d.get('uncertified_queries_last_step'), 'cpu', d.get('cpu_baseline',{}).get('value'))
except Exception as e:
    print('$f', 'ERR', e)
"; done
# ncu: launch list of the timed region (IVF f32 nprobe 32), then one full capture of the scan kernel
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/F_ivf_launches.csv python bench.py --workload ivf --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/F_ncu_ivf_l.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/F_flat_launches.csv python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/F_ncu_flat_l.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ivf_tc_kernel -c 1 -o gpurun_out/F_ivf_tc_f32 -f python bench.py --workload ivf --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/F_ncu_ivf_f.log 2>&1
ls -la gpurun_out/F_*
