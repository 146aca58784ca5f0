timeout 900 python bench.py --impl reference > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
timeout 900 python bench.py > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err
COALGPU_LIB=/root/repo/variants/r7.so timeout 900 ncu --set full --clock-control none -k regex:history_lockstep -s 30 -c 1 -o gpurun_out/prof_r2_ls6_r7 -f python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2f_ncu_r7.log 2>&1
for f in r2f_bench_ref r2f_bench_1gpu; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value'],1), (d.get('e2e') or {}), d.get('ms_per_step'))
except Exception as e: print('$f', 'FAILED', e)
PY
done
