timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "lockstep_shape" > gpurun_out/r2_pytest_s9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_s9.log; tail -12 gpurun_out/r2_pytest_s9.log
B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
COALGPU_LOCKSTEP_NH=96 $B > gpurun_out/r2_t_96.json 2> gpurun_out/r2_t_96.err
$B > gpurun_out/r2_t_def.json 2> gpurun_out/r2_t_def.err
for v in 96 def; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_t_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['value']), round(d['roofline']['kernel_ms'],1), d['details']['capacity'][:160])
except Exception as e: print('$v', 'FAILED', e)
"; done
