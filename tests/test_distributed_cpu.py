"""Host-side logic of the multi-GPU path on the CPU: the index sharding and the one-collective
(max, sum) combine, exercised with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from coalescenator_b200.distributed import combine_partials, merge_pairs, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for first, count in spans:
                assert first == pos
                pos += count
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _pairs_of(x):
    m = x.max(axis=0)
    return np.stack([m, np.exp(x - m).sum(axis=0)], axis=-1)


def test_merge_pairs_equals_global_logsumexp():
    rng = np.random.default_rng(0)
    x = rng.normal(-300, 40, size=(1000, 6, 5))
    full = _pairs_of(x)
    parts = [torch.from_numpy(_pairs_of(x[a:b])) for a, b in ((0, 10), (10, 400), (400, 1000))]
    empty = torch.zeros_like(parts[0]); empty[..., 0] = -float("inf")
    merged = merge_pairs(torch.stack(parts + [empty])).numpy()
    lse = merged[..., 0] + np.log(merged[..., 1])
    np.testing.assert_allclose(lse, full[..., 0] + np.log(full[..., 1]), rtol=1e-14)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    x = rng.normal(-250, 30, size=(501, 4, 3))
    first, count = shard_range(x.shape[0], rank, world)
    part = torch.from_numpy(_pairs_of(x[first:first + count]))
    total = combine_partials(part)
    q.put((rank, total.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_combine_partials_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(1)
    x = rng.normal(-250, 30, size=(501, 4, 3))
    full = _pairs_of(x)
    assert np.array_equal(got[0], got[1])          # identical on every rank
    lse = got[0][..., 0] + np.log(got[0][..., 1])
    np.testing.assert_allclose(lse, full[..., 0] + np.log(full[..., 1]), rtol=1e-14)


def test_shard_pairs_partitions():
    from coalescenator_b200.distributed import shard_pairs
    for n in (0, 1, 5, 26, 200):
        for world in (1, 2, 3, 8):
            owned = [shard_pairs(n, r, world) for r in range(world)]
            assert sorted(k for o in owned for k in o) == list(range(n))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def _pair_worker(rank, world, port, q):
    from coalescenator_b200.distributed import run_sampler_batch_distributed
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    problems = ["pair%d" % k for k in range(7)]
    seen = []

    def compute(ps):            # stands in for coalgpu_run_sampler_batch: a result that names its pair and its rank
        seen.extend(ps)
        return [dict(pair=p, rank=rank, grid=np.full((2, 3), float(p[4:]))) for p in ps]

    out = run_sampler_batch_distributed(problems, 100, seed=1, compute=compute)
    q.put((rank, seen, [(r["pair"], r["rank"], float(r["grid"].sum())) for r in out]))
    dist.barrier()
    dist.destroy_process_group()


def test_pairs_over_ranks_gloo_world2():
    """Locus pairs as the parallel axis: each rank computes only its own pairs, every rank ends
    up with all results in pair order."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pair_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {r: (seen, out) for r, seen, out in (q.get(timeout=120) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][0] == ["pair0", "pair2", "pair4", "pair6"] and got[1][0] == ["pair1", "pair3", "pair5"]
    assert got[0][1] == got[1][1]
    assert [x[0] for x in got[0][1]] == ["pair%d" % k for k in range(7)]
    assert [x[1] for x in got[0][1]] == [0, 1, 0, 1, 0, 1, 0]
    assert [x[2] for x in got[0][1]] == [6.0 * k for k in range(7)]
