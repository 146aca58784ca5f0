B="timeout 400 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
for c in cfg3 cfg4; do
$B --config $c > gpurun_out/r2_y_def_$c.json 2> gpurun_out/r2_y_def_$c.err
COALGPU_LIB=/root/repo/variants/notma6.so $B --config $c > gpurun_out/r2_y_notma6_$c.json 2> gpurun_out/r2_y_notma6_$c.err
done
for v in def_cfg3 notma6_cfg3 def_cfg4 notma6_cfg4; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_y_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['value']), round(d['roofline']['kernel_ms'],1))
except Exception as e: print('$v', 'FAILED', e)
"; done
COALGPU_LIB=/root/repo/variants/timing_notma.so timeout 300 python scripts_tmp/phase_clocks.py 16 2 11 20 > gpurun_out/r2_phase_clocks3.txt 2>&1; cat gpurun_out/r2_phase_clocks3.txt
