"""Many locus pairs with few samples each (the regime of a real analysis): end-to-end wall-clock
throughput of coalgpu_run_sampler_batch, host table construction included. Output: profiles/r1_batch_pipeline.json."""
import sys, time, json
sys.path.insert(0, '.')
import coalescenator_b200 as cb
from coalescenator_b200.problem import make_config, sub_problem
out = {}
smp = cb.ImportanceSampler()
for cfg, npairs, n in ((1, 26, 1000), (2, 200, 1000), (4, 200, 1000)):
    aln, loci, pairs = make_config(cfg)
    step = max(1, len(pairs) // npairs)
    probs = [sub_problem(aln, loci, *pairs[k]) for k in range(0, len(pairs), step)][:npairs]
    smp.run_sampler(probs[0], 64, seed=1)
    t0 = time.perf_counter(); rs = smp.run_sampler_batch(probs, n, seed=42); t1 = time.perf_counter()
    rs2 = smp.run_sampler_batch(probs, n, seed=42); t2 = time.perf_counter()
    t5 = time.perf_counter(); smp.run_sampler_batch(probs, 8 * n, seed=42); t6 = time.perf_counter()   # same pairs, 8x the samples: the device-bound rate of this pair mix
    k = min(len(probs), 16)
    t3 = time.perf_counter(); [smp.run_sampler(p, n, seed=42) for p in probs[:k]]; t4 = time.perf_counter()
    out['cfg%d' % cfg] = dict(pairs=len(probs), samples_per_pair=n, events_per_history=sum(r.n_events for r in rs) / (len(probs) * n),
                              batch_first_call_hist_per_s=len(probs) * n / (t1 - t0), batch_second_call_hist_per_s=len(probs) * n / (t2 - t1),
                              one_call_per_pair_hist_per_s=k * n / (t4 - t3),
                              batch_8x_samples_hist_per_s=len(probs) * 8 * n / (t6 - t5))
print(json.dumps(out, indent=1))
