B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
COALGPU_LOCKSTEP_NH=32 COALGPU_LOCKSTEP_NT=128 $B > gpurun_out/r2_s_32128.json 2> gpurun_out/r2_s_32128.err
COALGPU_LOCKSTEP_NH=32 COALGPU_LOCKSTEP_NT=64 $B > gpurun_out/r2_s_3264.json 2> gpurun_out/r2_s_3264.err
COALGPU_LOCKSTEP_NH=8 $B > gpurun_out/r2_s_832.json 2> gpurun_out/r2_s_832.err
COALGPU_LOCKSTEP_NH=16 COALGPU_LOCKSTEP_NT=128 $B > gpurun_out/r2_s_16128.json 2> gpurun_out/r2_s_16128.err
COALGPU_LIB=/root/repo/variants/r7.so $B > gpurun_out/r2_s_r7.json 2> gpurun_out/r2_s_r7.err
for v in 32128 3264 832 16128 r7; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_s_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['value']), round(d['roofline']['kernel_ms'],1))
except Exception as e: print('$v', 'FAILED', e)
"; done
