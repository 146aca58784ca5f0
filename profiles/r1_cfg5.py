"""BASELINE config 5: ONE locus pair of the cfg2 data, N = 1e8 importance samples over all
ranks (torchrun, one process per GPU), NCCL combine of the (max, sum) partials at the end. Rank 0 prints one JSON line.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 profiles/r1_cfg5.py [N]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from coalescenator_b200.distributed import run_sampler_distributed
from coalescenator_b200.problem import make_config, sub_problem

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
world = int(os.environ.get("WORLD_SIZE", 1))
if world > 1:
    sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local)); dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
aln, loci, pairs = make_config(2)
p = sub_problem(aln, loci, *pairs[40])
run_sampler_distributed(p, 200_000 * world, seed=1, device=local)          # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
out = run_sampler_distributed(p, n, seed=42, device=local, chunk=1 << 22)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
if int(os.environ.get("RANK", 0)) == 0:
    print(json.dumps(dict(config="cfg5: one cfg2 locus pair (LA=%d, LB=%d), N=%d" % (p.LA, p.LB, n), n_gpus=world, seconds=dt, histories_per_s=n / dt,
                          log_sum_w=float(out.likelihood.ravel()[0]), ess=float(out.ess.ravel()[0]), events_per_history=out.n_events / n)))
if world > 1:
    dist.destroy_process_group()
