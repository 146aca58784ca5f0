B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2_x_def.json 2> gpurun_out/r2_x_def.err
for v in notma6 notma7 notma8 notma9; do COALGPU_LIB=/root/repo/variants/$v.so $B > gpurun_out/r2_x_$v.json 2> gpurun_out/r2_x_$v.err; done
for v in def notma6 notma7 notma8 notma9; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_x_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['value']), round(d['roofline']['kernel_ms'],1))
except Exception as e: print('$v', 'FAILED', e)
"; done
