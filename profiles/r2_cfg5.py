"""BASELINE config 5: ONE locus pair of the cfg2 data, N = 1e8 importance samples sharded over the
visible GPUs INSIDE libcoalgpu.so (coalgpu_run_sampler_multi: one host thread per device, per-device reduction, NCCL
allreduce MAX + rescale + SUM over the library's own communicator). One process; prints one JSON line.
  python profiles/r2_cfg5.py [N] [n_gpus]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coalescenator_b200 as cb
from coalescenator_b200.problem import make_config, sub_problem
from coalescenator_b200.sampler import nccl_calls

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
ngpu = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
aln, loci, pairs = make_config(2)
p = sub_problem(aln, loci, *pairs[40])
S = cb.ImportanceSampler()
devs = list(range(ngpu))
S.run_sampler_multi(p, 200_000 * ngpu, devices=devs, seed=1)          # warm-up: communicator, kernels, table cache
c0 = nccl_calls()
t0 = time.perf_counter()
out = S.run_sampler_multi(p, n, devices=devs, seed=42)
dt = time.perf_counter() - t0
print(json.dumps(dict(config="cfg5: one cfg2 locus pair (LA=%d, LB=%d), N=%d, sample-sharded inside libcoalgpu.so" % (p.LA, p.LB, n), n_gpus=ngpu,
                      seconds=dt, histories_per_s=n / dt, device_seconds_max=out.seconds_device, nccl_collectives=nccl_calls() - c0,
                      log_sum_w=float(out.likelihood.ravel()[0]), log_sum_w2=float(out.likelihoodSq.ravel()[0]), ess=float(out.ess.ravel()[0]),
                      events_per_history=out.n_events / n)))
