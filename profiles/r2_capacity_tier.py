"""Throughput of the capacity tier (type store in global memory) on the stress reading of BASELINE cfg4: 10 sampling times x 200
sequences = 2,000 lineages per locus pair (the problem of tests/test_gpu_parity.py::test_capacity_tier_type_store_in_global_memory).
Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coalescenator_b200 as cb
from coalescenator_b200.problem import synth_alignment, tile_loci, locus_pairs, sub_problem
aln = synth_alignment(3, 10, 200, 4000, 120)
loci = tile_loci(aln, 100)
p = sub_problem(aln, loci, *locus_pairs(loci)[3])
S = cb.ImportanceSampler()
S.run_sampler(p, 64, seed=1)
out = {}
for n in (2048, 16384):
    t0 = time.perf_counter(); r = S.run_sampler(p, n, seed=7); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out[str(n)] = dict(seconds=dt, histories_per_s=n / dt, device_seconds=r.seconds_device, events_per_history=r.n_events / n)
s = cb.DeviceSampler(p, device=0)
b = s.alloc(2048); s.sample(b, 7, 0, 2048); torch.cuda.synchronize()
print(json.dumps(dict(problem="10 sampling times x 200 sequences, LA=%d LB=%d, %d lineages" % (p.LA, p.LB, p.n_lineages()), runs=out, capacity=s.counters())))
