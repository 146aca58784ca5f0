timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_full8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_full8.log; tail -8 gpurun_out/r2_pytest_full8.log
B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
for v in r8 notma notma_r8; do COALGPU_LIB=/root/repo/variants/$v.so $B > gpurun_out/r2_h_$v.json 2> gpurun_out/r2_h_$v.err; done
for v in r8 notma notma_r8; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_h_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['value']), round(d['roofline']['kernel_ms'],1))
except Exception as e: print('$v', 'FAILED', e)
"; done
timeout 900 python profiles/r2_cfg5.py 1e8 1 > gpurun_out/r2_cfg5_1gpu.json 2> gpurun_out/r2_cfg5_1gpu.err; cat gpurun_out/r2_cfg5_1gpu.json
