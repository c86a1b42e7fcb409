timeout 900 python -m pytest tests/test_gpu_ivf.py -x -q -m gpu > gpurun_out/T4_pytest.log 2>&1; echo "pytest rc $?"; tail -15 gpurun_out/T4_pytest.log
for i in 1 2; do for o in 0 1; do timeout 200 python bench.py --workload flat --no-cpu-baseline --steps 20 --warmup 5 --option tc_strided=$o 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('flat strided=$o', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['clocks'])"; done; done
