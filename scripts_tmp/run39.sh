timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2_pytest_s8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_s8.log; tail -15 gpurun_out/r2_pytest_s8.log
B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2_q3_def.json 2> gpurun_out/r2_q3_def.err
python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_q3_def.json').read().strip().splitlines()[-1])
    print('def', round(d['value']), round(d['roofline']['kernel_ms'],1))
except Exception as e: print('def', 'FAILED', e)
"
COALGPU_LIB=/root/repo/variants/timing.so timeout 300 python scripts_tmp/phase_clocks.py 16 2 11 20 > gpurun_out/r2_phase_clocks6.txt 2>&1; cat gpurun_out/r2_phase_clocks6.txt
