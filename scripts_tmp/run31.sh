timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "lockstep_shape or capacity_tier" > gpurun_out/r2_pytest_s3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_s3.log; tail -25 gpurun_out/r2_pytest_s3.log
B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2_v_def.json 2> gpurun_out/r2_v_def.err
COALGPU_LOCKSTEP_NH=8 $B > gpurun_out/r2_v_832.json 2> gpurun_out/r2_v_832.err
COALGPU_LOCKSTEP_NH=16 COALGPU_LOCKSTEP_NT=128 $B > gpurun_out/r2_v_16128.json 2> gpurun_out/r2_v_16128.err
for v in def 832 16128; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_v_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['value']), round(d['roofline']['kernel_ms'],1), d['details']['capacity'][:200])
except Exception as e: print('$v', 'FAILED', e)
"; done
