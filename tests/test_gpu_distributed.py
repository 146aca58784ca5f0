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


def test_native_multi_entry_point_on_one_device(cuda_lib):
    """coalgpu_run_sampler_multi with a single device (no NCCL involved): the same sharding / chunk / merge code as the
    multi-device call must return what the drop-in call returns, bit for bit."""
    import coalescenator_b200 as cb
    p, _, _ = load_golden("stat_grid")
    one = cb.ImportanceSampler().run_sampler(p, 3001, seed=9)
    out = cb.ImportanceSampler().run_sampler_multi(p, 3001, devices=[0], seed=9)
    assert np.array_equal(out.likelihood, one.likelihood) and np.array_equal(out.likelihoodSq, one.likelihoodSq)
    assert np.array_equal(out.tmrca, one.tmrca) and np.array_equal(out.lineages_remaining, one.lineages_remaining)
    assert (out.n_events, out.n_pismc, out.n_histories) == (one.n_events, one.n_pismc, 3001)


def test_native_multi_gpu_nccl_equals_single_gpu(cuda_lib):
    """The C-ABI sample-sharded path (VERDICT r1 task 2): one host thread per device inside libcoalgpu.so, per-device
    reduction, ncclAllReduce(MAX) + rescale + ncclAllReduce(SUM) over the library's own communicator. Grids for
    2, 3, ... all visible devices (incl. a count that does not divide the sample count) equal the single-device call to
    1e-13; event and pi_SMC counts are identical; NCCL collectives were really issued by the library."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs 2 GPUs")
    import coalescenator_b200 as cb
    from coalescenator_b200.sampler import nccl_calls
    for name, n in (("stat_grid", 3001), ("cfg2_pair0", 20000)):
        p, _, _ = load_golden(name)
        one = cb.ImportanceSampler().run_sampler(p, n, seed=11)
        for R in sorted({2, 3, ndev} & set(range(2, ndev + 1))):
            before = nccl_calls()
            out = cb.ImportanceSampler().run_sampler_multi(p, n, devices=list(range(R)), seed=11)
            assert nccl_calls() - before == 2 * R          # one MAX and one SUM allreduce per device
            np.testing.assert_allclose(out.likelihood, one.likelihood, rtol=1e-13)
            np.testing.assert_allclose(out.likelihoodSq, one.likelihoodSq, rtol=1e-13)
            np.testing.assert_allclose(out.tmrca, one.tmrca, rtol=1e-11, atol=1e-13)
            np.testing.assert_allclose(out.lineages_remaining, one.lineages_remaining, rtol=1e-11, atol=1e-13)
            assert (out.n_events, out.n_pismc) == (one.n_events, one.n_pismc)
    # fewer samples than devices: the ranks without a history carry (-inf, 0) through both allreduces
    p, _, _ = load_golden("cfg1_pair0")
    one = cb.ImportanceSampler().run_sampler(p, 1, seed=3)
    out = cb.ImportanceSampler().run_sampler_multi(p, 1, devices=list(range(ndev)), seed=3)
    np.testing.assert_allclose(out.likelihood, one.likelihood, rtol=1e-13)
