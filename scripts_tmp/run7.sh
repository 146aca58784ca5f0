timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_ls2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_ls2.log; tail -4 gpurun_out/r2_pytest_ls2.log
B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2_c_def.json 2> gpurun_out/r2_c_def.err
COALGPU_LOCKSTEP_NH=32 COALGPU_LOCKSTEP_NT=64 $B > gpurun_out/r2_c_3264.json 2> gpurun_out/r2_c_3264.err
COALGPU_LOCKSTEP_NH=32 COALGPU_LOCKSTEP_NT=128 $B > gpurun_out/r2_c_32128.json 2> gpurun_out/r2_c_32128.err
for v in r8 notma notma_r8; do COALGPU_LIB=/root/repo/variants/$v.so $B > gpurun_out/r2_c_$v.json 2> gpurun_out/r2_c_$v.err; done
COALGPU_LIB=/root/repo/variants/notma_r8.so COALGPU_LOCKSTEP_NH=32 COALGPU_LOCKSTEP_NT=64 $B > gpurun_out/r2_c_notma_r8_3264.json 2> gpurun_out/r2_c_notma_r8_3264.err
COALGPU_LIB=/root/repo/variants/r8.so COALGPU_LOCKSTEP_NH=32 COALGPU_LOCKSTEP_NT=64 $B > gpurun_out/r2_c_r8_3264.json 2> gpurun_out/r2_c_r8_3264.err
for v in def 3264 32128 r8 notma notma_r8 notma_r8_3264 r8_3264; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_c_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['value']), round(d['roofline']['kernel_ms'],1))
except Exception as e: print('$v', 'FAILED', e)
"; done
