run() { timeout 200 python bench.py --workload $1 --no-cpu-baseline --steps 20 --warmup 5 $2 $3 $4 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$1 $2 $3 $4', 'ms', round(d['ms_per_step'],3), 'kern', round(d['roofline']['kernel_ms'],3), 'qps', round(d['value']), 'parity', d['parity_sample']['ids_equal'], d['parity_sample']['dist_bits_equal'], 'uncert', d['uncertified_queries_last_step'], d['clocks']['sm_mhz'])"; }
run flat
run flat --option tc_ts=0
run flat --option tc_f32_lo_smem=1
run c5
run c5 --option tc_ts=0
run c5 --option tc_f32_lo_smem=1
run flat --metric euclidean
