set -x
timeout 300 python bench.py > gpurun_out/A_flat_f32.json 2> gpurun_out/A_flat_f32.err
timeout 200 python bench.py --dtype bf16 --no-cpu-baseline > gpurun_out/A_flat_bf16.json 2> gpurun_out/A_flat_bf16.err
timeout 200 python bench.py --dtype sq8 --no-cpu-baseline > gpurun_out/A_flat_sq8.json 2> gpurun_out/A_flat_sq8.err
for np in 8 32 128; do timeout 300 python bench.py --workload ivf --nprobe $np --recall --no-cpu-baseline > gpurun_out/A_ivf_f32_np$np.json 2> gpurun_out/A_ivf_f32_np$np.err; done
for dt in bf16 sq8; do timeout 300 python bench.py --workload ivf --dtype $dt --recall --no-cpu-baseline > gpurun_out/A_ivf_$dt.json 2> gpurun_out/A_ivf_$dt.err; done
for f in gpurun_out/A_*.json; do python -c "
import json,sys
d=json.load(open('$f'))
r=d['roofline']
print('$f'.split('/')[-1], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'kern_ms', round(r['kernel_ms'],3), 'frac', round(r['frac'],3), d.get('recall_at_k_vs_exact_f32',{}).get('value'), d.get('parity_sample'), d.get('cpu_baseline',{}).get('value'))
"; done
