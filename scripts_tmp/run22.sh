B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2_u_def.json 2> gpurun_out/r2_u_def.err
for v in s1q2 s4q2 s2q1 s2q4 s4q4; do COALGPU_LIB=/root/repo/variants/$v.so $B > gpurun_out/r2_u_$v.json 2> gpurun_out/r2_u_$v.err; done
for v in def s1q2 s4q2 s2q1 s2q4 s4q4; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_u_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['value']), round(d['roofline']['kernel_ms'],1))
except Exception as e: print('$v', 'FAILED', e)
"; done
