timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2_pytest_s10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_s10.log; tail -12 gpurun_out/r2_pytest_s10.log
B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2_w2_def.json 2> gpurun_out/r2_w2_def.err
COALGPU_LIB=/root/repo/variants/tma.so $B > gpurun_out/r2_w2_tma.json 2> gpurun_out/r2_w2_tma.err
for v in def tma; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_w2_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['value']), round(d['roofline']['kernel_ms'],1))
except Exception as e: print('$v', 'FAILED', e)
"; done
COALGPU_LIB=/root/repo/variants/timing.so timeout 300 python scripts_tmp/phase_clocks.py 16 2 11 20 > gpurun_out/r2_phase_clocks7.txt 2>&1; cat gpurun_out/r2_phase_clocks7.txt
COALGPU_LIB=/root/repo/variants/tma.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "history_events or random_problems or two_pass or lockstep_shape or capacity_tier" > gpurun_out/r2_pytest_tma_build.log 2>&1; echo "pytest rc=$? (libcoalgpu built with -DCOAL_TMA)" >> gpurun_out/r2_pytest_tma_build.log; tail -4 gpurun_out/r2_pytest_tma_build.log
