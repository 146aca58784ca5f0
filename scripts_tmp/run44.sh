timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_e2e_a.json 2> gpurun_out/r2_e2e_a.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_e2e_a.json').read().strip().splitlines()[-1])
print('def', round(d['value']), round(d['roofline']['kernel_ms'],1), d['e2e'])"
COALGPU_TIMING=1 timeout 300 python - <<'PY' 2>&1 | tail -12
import sys; sys.path.insert(0,'/root/repo')
import bench, coalescenator_b200 as cb, time, torch
probs = bench.bench_problems()
for p in probs[:3]:
    for rep in range(2):
        t=time.perf_counter(); r = cb.ImportanceSampler().run_sampler(p, 98304, seed=1); torch.cuda.synchronize(); print('pair', p.LA, p.LB, 'wall %.1f ms device %.1f ms'%(1e3*(time.perf_counter()-t), 1e3*r.seconds_device))
PY
