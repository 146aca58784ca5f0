B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
COALGPU_CARVEOUT=100 $B > gpurun_out/r2_a_c100.json 2> gpurun_out/r2_a_c100.err
COALGPU_CARVEOUT=75 $B > gpurun_out/r2_a_c75.json 2> gpurun_out/r2_a_c75.err
COALGPU_CARVEOUT=60 $B > gpurun_out/r2_a_c60.json 2> gpurun_out/r2_a_c60.err
for v in u11 u14 r7 r8; do COALGPU_LIB=/root/repo/variants/$v.so $B > gpurun_out/r2_a_$v.json 2> gpurun_out/r2_a_$v.err; done
COALGPU_CARVEOUT=100 COALGPU_LIB=/root/repo/variants/r8.so $B > gpurun_out/r2_a_r8c100.json 2> gpurun_out/r2_a_r8c100.err
for v in c100 c75 c60 u11 u14 r7 r8 r8c100; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_a_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['value']), round(d['roofline']['kernel_ms'],1))
except Exception as e: print('$v', 'FAILED', e)
"; done
