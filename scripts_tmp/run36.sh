timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_s5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_s5.log; tail -5 gpurun_out/r2_pytest_s5.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_z_def.json 2> gpurun_out/r2_z_def.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_z_def.json').read().strip().splitlines()[-1])
print('def', round(d['value']), round(d['roofline']['kernel_ms'],1), d['e2e'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:history_lockstep -s 30 -c 1 -o gpurun_out/prof_r2_ls4 -f python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_ls4.log 2>&1; tail -3 gpurun_out/r2_ncu_ls4.log
