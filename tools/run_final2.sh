set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/I_pytest.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/I_pytest.log
for np in 8 32 128; do timeout 200 python bench.py --workload ivf --nprobe $np --recall --no-cpu-baseline > gpurun_out/I_ivf_f32_np$np.json 2> /dev/null; done
for dt in bf16 sq8; do timeout 200 python bench.py --workload ivf --dtype $dt --recall --no-cpu-baseline > gpurun_out/I_ivf_$dt.json 2> /dev/null; done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/I_ivf_launches.csv python bench.py --workload ivf --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/I_ncu_ivf_l.log 2>&1
for f in gpurun_out/I_*.json; do python -c "
import json,sys
d=json.load(open('$f')); r=d.get('roofline',{})
print('$f'.split('/')[-1], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'kern_ms', round(r.get('kernel_ms',0),3), 'frac', round(r.get('frac',0),3), d.get('recall_at_k_vs_exact_f32',{}).get('value'), 'uncert', d.get('uncertified_queries_last_step'))
"; done
