timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_1gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_1gpu.log; tail -4 gpurun_out/r2f_pytest_1gpu.log
timeout 900 python bench.py --impl reference > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
timeout 900 python bench.py > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err
timeout 600 python bench.py --config cfg3 --no-cpu-baseline > gpurun_out/r2f_bench_cfg3.json 2> gpurun_out/r2f_bench_cfg3.err
timeout 600 python bench.py --config cfg4 --no-cpu-baseline > gpurun_out/r2f_bench_cfg4.json 2> gpurun_out/r2f_bench_cfg4.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2f_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:history_lockstep -s 30 -c 1 -o gpurun_out/prof_r2_ls6 -f python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2f_ncu_ls6.log 2>&1
timeout 600 python profiles/r2_cfg5.py 1e8 1 > gpurun_out/r2f_cfg5_1gpu.json 2> gpurun_out/r2f_cfg5_1gpu.err; cat gpurun_out/r2f_cfg5_1gpu.json
for f in r2f_bench_ref r2f_bench_1gpu r2f_bench_cfg3 r2f_bench_cfg4; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value'],1), (d.get('e2e') or {}).get('value'), d.get('ms_per_step'))
except Exception as e: print('$f', 'FAILED', e)
PY
done
