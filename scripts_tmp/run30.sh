timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_s2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_s2.log; tail -4 gpurun_out/r2_pytest_s2.log
COALGPU_LIB=/root/repo/variants/timing.so timeout 300 python scripts_tmp/phase_clocks.py 16 2 11 0 5 20 > gpurun_out/r2_phase_clocks.txt 2>&1; cat gpurun_out/r2_phase_clocks.txt
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_ref_default.json 2> gpurun_out/r2_bench_ref_default.err
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_default.json 2> gpurun_out/r2_bench_1gpu_default.err
timeout 600 python bench.py --config cfg3 --no-cpu-baseline > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err
timeout 600 python bench.py --config cfg4 --no-cpu-baseline > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_ls.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_launches.log 2>&1
for f in r2_bench_ref_default r2_bench_1gpu_default r2_bench_cfg3 r2_bench_cfg4; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value'],1), d.get('e2e'), d.get('ms_per_step'))
except Exception as e: print('$f', 'FAILED', e)
PY
done
