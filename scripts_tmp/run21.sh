timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_quad.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_quad.log; tail -12 gpurun_out/r2_pytest_quad.log
B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2_q_def.json 2> gpurun_out/r2_q_def.err
COALGPU_LOCKSTEP_NH=32 $B > gpurun_out/r2_q_32128.json 2> gpurun_out/r2_q_32128.err
for v in def 32128; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_q_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['value']), round(d['roofline']['kernel_ms'],1))
except Exception as e: print('$v', 'FAILED', e)
"; done
COALGPU_LIB=/root/repo/variants/timing.so python scripts_tmp/phase_clocks.py 16 2 11
