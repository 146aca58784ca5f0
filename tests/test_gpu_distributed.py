"""Multi-GPU path on real devices: run_sampler_distributed over NCCL must return, on every rank, the same grids as
the single-GPU drop-in call (histories are a pure function of (seed, index); the ranks' (max, sum) partials are
merged in rank order). Needs >= 2 GPUs for the NCCL leg; the single-process leg runs on one."""
import os
import socket

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_distributed_entry_point_single_process(cuda_lib):
    import coalescenator_b200 as cb
    from coalescenator_b200.distributed import run_sampler_distributed
    p, _, _ = load_golden("stat_grid")
    one = cb.ImportanceSampler().run_sampler(p, 3000, seed=9)
    out = run_sampler_distributed(p, 3000, seed=9, device=0, chunk=1024)     # 3 chunks merged on the device
    np.testing.assert_allclose(out.likelihood, one.likelihood, rtol=1e-13)
    np.testing.assert_allclose(out.likelihoodSq, one.likelihoodSq, rtol=1e-13)
    np.testing.assert_allclose(out.tmrca, one.tmrca, rtol=1e-11, atol=1e-13)
    assert out.n_events == one.n_events


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from coalescenator_b200.distributed import run_sampler_distributed
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    p, _, _ = load_golden("stat_grid")
    out = run_sampler_distributed(p, 3001, seed=9, device=rank)
    q.put((rank, out.likelihood.copy(), out.likelihoodSq.copy(), out.n_events))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpus_nccl_equal_single_gpu(cuda_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import coalescenator_b200 as cb
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = dict((r, (a, b, n)) for r, a, b, n in (q.get(timeout=300) for _ in range(2)))
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    p, _, _ = load_golden("stat_grid")
    one = cb.ImportanceSampler().run_sampler(p, 3001, seed=9)
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])     # identical on both ranks
    np.testing.assert_allclose(got[0][0], one.likelihood, rtol=1e-13)
    np.testing.assert_allclose(got[0][1], one.likelihoodSq, rtol=1e-13)
    assert got[0][2] == one.n_events


def _pair_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from coalescenator_b200.distributed import run_sampler_batch_distributed
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    probs = [load_golden(n)[0] for n in ("cfg1_pair0", "cfg2_pair0", "missing_pair3", "stat_grid", "varV_req_n")]
    out = run_sampler_batch_distributed(probs, 400, seed=4, device=rank)
    q.put((rank, [o.likelihood.copy() for o in out], [o.n_events for o in out]))
    dist.barrier()
    dist.destroy_process_group()


def test_pairs_over_two_gpus_equal_single_gpu_batch(cuda_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import coalescenator_b200 as cb
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pair_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = {r: (a, n) for r, a, n in (q.get(timeout=300) for _ in range(2))}
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    probs = [load_golden(n)[0] for n in ("cfg1_pair0", "cfg2_pair0", "missing_pair3", "stat_grid", "varV_req_n")]
    one = cb.ImportanceSampler().run_sampler_batch(probs, 400, seed=4)
    for k in range(len(probs)):
        assert np.array_equal(got[0][0][k], got[1][0][k]) and np.array_equal(got[0][0][k], one[k].likelihood)
        assert got[0][1][k] == one[k].n_events
