"""One-off throughput numbers for the other BASELINE configs (recorded in profiles/r1_config_throughput.md)."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
import coalescenator_b200 as cb
from coalescenator_b200.problem import make_config, sub_problem, grid_cfg3
out = {}
smp = cb.ImportanceSampler()
# cfg1: 26 pairs x 1000 samples, sequential vs batched
aln, loci, pairs = make_config(1)
probs = [sub_problem(aln, loci, *pr) for pr in pairs]
smp.run_sampler(probs[0], 100, seed=1)
t0 = time.perf_counter(); [smp.run_sampler(p, 1000, seed=42) for p in probs]; t1 = time.perf_counter()
smp.run_sampler_batch(probs, 1000, seed=42); t2 = time.perf_counter()
out['cfg1'] = dict(pairs=len(probs), samples_per_pair=1000, sequential_hist_per_s=len(probs) * 1000 / (t1 - t0), batched_hist_per_s=len(probs) * 1000 / (t2 - t1))
# cfg3: G = 1024, 1e4 samples, one pair
aln, loci, pairs = make_config(2)
mu, rho, lam = grid_cfg3()
p3 = sub_problem(aln, loci, *pairs[100], target_mu=mu, target_rho=rho, target_lambda=lam)
smp.run_sampler(p3, 1000, seed=1)
t0 = time.perf_counter(); r = smp.run_sampler(p3, 10000, seed=42); t1 = time.perf_counter()
out['cfg3'] = dict(G=p3.G, LA=p3.LA, LB=p3.LB, samples=10000, e2e_hist_per_s=10000 / (t1 - t0), device_s=r.seconds_device, events_per_history=r.n_events / 10000)
p2 = sub_problem(aln, loci, *pairs[100])
t0 = time.perf_counter(); r = smp.run_sampler(p2, 10000, seed=42); t1 = time.perf_counter()
out['cfg2_same_pair_G1'] = dict(samples=10000, e2e_hist_per_s=10000 / (t1 - t0), device_s=r.seconds_device)
# cfg4: 10 time points x 20 sequences, 100 kb: 8 pairs x 4096 samples
aln, loci, pairs = make_config(4)
probs = [sub_problem(aln, loci, *pairs[k]) for k in range(0, 3200, 400)]
smp.run_sampler(probs[0], 256, seed=1)
t0 = time.perf_counter(); rs = [smp.run_sampler(p, 4096, seed=42) for p in probs]; t1 = time.perf_counter()
out['cfg4'] = dict(pairs=len(probs), LAB=[(p.LA, p.LB) for p in probs], samples_per_pair=4096, e2e_hist_per_s=len(probs) * 4096 / (t1 - t0),
                   device_hist_per_s=len(probs) * 4096 / sum(r.seconds_device for r in rs), events_per_history=sum(r.n_events for r in rs) / (len(probs) * 4096))
print(json.dumps(out, indent=1))
