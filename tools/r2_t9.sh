timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_flat.py -x -q -m gpu > gpurun_out/T9_pytest.log 2>&1; echo "pytest rc $?"; tail -25 gpurun_out/T9_pytest.log
for o in 1 0; do timeout 200 python bench.py --workload flat --no-cpu-baseline --steps 20 --warmup 5 --option tc_f32_fp16=$o 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('flat fp16=$o', 'ms', round(d['ms_per_step'],3), 'kern', round(d['roofline']['kernel_ms'],3), 'qps', round(d['value']), 'parity', d['parity_sample']['ids_equal'], d['parity_sample']['dist_bits_equal'], 'uncert', d['uncertified_queries_last_step'], 'fallback', d['fallback_queries_total'], d['clocks']['sm_mhz'])"; done
timeout 200 python bench.py --workload c5 --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('c5', 'ms', round(d['ms_per_step'],3), 'kern', round(d['roofline']['kernel_ms'],3), 'qps', round(d['value']), 'parity', d['parity_sample']['ids_equal'], d['parity_sample']['dist_bits_equal'], 'uncert', d['uncertified_queries_last_step'], 'fallback', d['fallback_queries_total'])"
