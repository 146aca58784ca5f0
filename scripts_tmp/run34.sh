B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
for v in notma5 notma4; do COALGPU_LIB=/root/repo/variants/$v.so $B > gpurun_out/r2_x_$v.json 2> gpurun_out/r2_x_$v.err; done
COALGPU_LOCKSTEP_NH=8 COALGPU_LIB=/root/repo/variants/notma6.so $B > gpurun_out/r2_x_notma6_832.json 2> gpurun_out/r2_x_notma6_832.err
COALGPU_LOCKSTEP_NH=32 COALGPU_LOCKSTEP_NT=128 COALGPU_LIB=/root/repo/variants/notma6.so $B > gpurun_out/r2_x_notma6_32128.json 2> gpurun_out/r2_x_notma6_32128.err
COALGPU_LOCKSTEP_NH=32 COALGPU_LOCKSTEP_NT=64 COALGPU_LIB=/root/repo/variants/notma6.so $B > gpurun_out/r2_x_notma6_3264.json 2> gpurun_out/r2_x_notma6_3264.err
COALGPU_LOCKSTEP_NH=16 COALGPU_LOCKSTEP_NT=128 COALGPU_LIB=/root/repo/variants/notma6.so $B > gpurun_out/r2_x_notma6_16128.json 2> gpurun_out/r2_x_notma6_16128.err
for v in notma5 notma4 notma6_832 notma6_32128 notma6_3264 notma6_16128; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_x_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['value']), round(d['roofline']['kernel_ms'],1))
except Exception as e: print('$v', 'FAILED', e)
"; done
