cat > /tmp/bp2.py <<'PY'
import sys, time, json
sys.path.insert(0, '/root/repo')
import coalescenator_b200 as cb
from coalescenator_b200.problem import make_config, sub_problem
smp = cb.ImportanceSampler()
aln, loci, pairs = make_config(2)
step = max(1, len(pairs) // 200)
probs = [sub_problem(aln, loci, *pairs[k]) for k in range(0, len(pairs), step)][:200]
smp.run_sampler(probs[0], 64, seed=1)
smp.run_sampler_batch(probs, 1000, seed=42)
t1 = time.perf_counter(); smp.run_sampler_batch(probs, 1000, seed=42); t2 = time.perf_counter()
print("cfg2 200 pairs x 1000: %.0f hist/s" % (200 * 1000 / (t2 - t1)))
t1 = time.perf_counter(); smp.run_sampler_batch(probs, 4000, seed=42); t2 = time.perf_counter()
print("cfg2 200 pairs x 4000: %.0f hist/s" % (200 * 4000 / (t2 - t1)))
PY
echo default; COALGPU_TIMING=1 python /tmp/bp2.py 2>&1 | tail -4
echo conn32; CUDA_DEVICE_MAX_CONNECTIONS=32 COALGPU_TIMING=1 python /tmp/bp2.py 2>&1 | tail -4
