timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_ls6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_ls6.log; tail -12 gpurun_out/r2_pytest_ls6.log
B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2_g_def.json 2> gpurun_out/r2_g_def.err
for v in def; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_g_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['value']), round(d['roofline']['kernel_ms'],1))
except Exception as e: print('$v', 'FAILED', e)
"; done
COALGPU_LIB=/root/repo/variants/timing.so python scripts_tmp/phase_clocks.py 16 2 11
