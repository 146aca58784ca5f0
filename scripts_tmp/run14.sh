B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
for v in r8 notma notma_r8 r7; do COALGPU_LIB=/root/repo/variants/$v.so $B > gpurun_out/r2_h_$v.json 2> gpurun_out/r2_h_$v.err; done
COALGPU_LOCKSTEP_NH=32 COALGPU_LOCKSTEP_NT=128 $B > gpurun_out/r2_h_32128.json 2> gpurun_out/r2_h_32128.err
timeout 600 python bench.py --config cfg3 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_h_cfg3.json 2> gpurun_out/r2_h_cfg3.err
timeout 600 python bench.py --config cfg4 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_h_cfg4.json 2> gpurun_out/r2_h_cfg4.err
for v in r8 notma notma_r8 r7 32128 cfg3 cfg4; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_h_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['value']), round(d['roofline']['kernel_ms'],1), d.get('e2e',{}).get('value'), d['details']['capacity'][-160:])
except Exception as e: print('$v', 'FAILED', e)
"; done
