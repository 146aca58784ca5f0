"""Multi-GPU sharding: the importance samples of one locus pair over the ranks
(run_sampler_distributed), or the locus pairs of an analysis over the ranks (run_sampler_batch_distributed).

Histories are i.i.d. given (data, driving values) and history h is a pure function of (seed, h)
(Philox counter), so rank r simply draws the contiguous index range shard_range(n, r, R) on its own
GPU with no data-path communication. The only exchange is at the end: every rank holds, per grid point
and statistic, a partial log-sum-exp as a (max, sum exp(x - max)) pair; ONE all_gather over NCCL moves
the R small pair arrays and each rank combines them locally, in rank order, so the result is
bit-identical on every rank and independent of arrival order. Reference analogue: the mutex-guarded
multiset merge ImportanceSampler.cpp:104-106 followed by sumSamples (SamplerOutput.cpp:115-133)."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .problem import Problem


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, first+count) of rank `rank`: floor(n/world) each, the remainder to the lowest ranks
    (the reference gives it to its first thread, ImportanceSampler.cpp:113-114,128)."""
    base, rem = divmod(int(n), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def merge_pairs(pairs: torch.Tensor) -> torch.Tensor:
    """pairs[R, ..., 2] of (max, sum) -> merged [..., 2]. Ranks without samples carry (-inf, 0)."""
    mx = pairs[..., 0]
    sm = pairs[..., 1]
    m = mx.max(dim=0).values
    scale = torch.exp(mx - m.unsqueeze(0))
    scale = torch.where(torch.isfinite(mx), scale, torch.zeros_like(scale))
    s = (sm * scale).sum(dim=0)
    return torch.stack([m, s], dim=-1)


def combine_partials(partials: torch.Tensor, group=None) -> torch.Tensor:
    """One collective: all_gather the per-rank (max, sum) pairs, merge locally in rank order."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return partials
    world = dist.get_world_size(group)
    parts = [torch.empty_like(partials) for _ in range(world)]
    dist.all_gather(parts, partials.contiguous(), group=group)
    return merge_pairs(torch.stack(parts))


def empty_partials(n_stats: int, G: int, device) -> torch.Tensor:
    p = torch.zeros((n_stats, G, 2), dtype=torch.float64, device=device)
    p[..., 0] = -float("inf")
    return p


def run_sampler_distributed(problem: Problem, num_iterations: int, seed: int = 0, quadrature_points: int = 16,
                            device: Optional[int] = None, group=None, chunk: int = 1 << 20):
    """runSampler over all ranks of the process group (one process per GPU). Every rank returns the
    same SamplerOutput. Requires torch.distributed to be initialised with the nccl backend."""
    from .sampler import DeviceSampler
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(device)
    first, count = shard_range(num_iterations, rank, world)
    s = DeviceSampler(problem, quadrature_points, device)
    acc = empty_partials(3 + problem.T, problem.G, torch.device("cuda", device))
    buf = s.alloc(max(1, min(count, chunk)))
    done = 0
    while done < count:
        n = min(chunk, count - done)
        s.sample(buf, seed, first + done, n)
        part = s.partials(buf, n)
        acc = merge_pairs(torch.stack([acc, part]))
        if done == 0 and count > n:
            s.tune()          # size the first-pass type store / lanes per history from what the first chunk needed
        done += n
    total = combine_partials(acc, group)
    out = s.finalize(total.cpu().numpy(), num_iterations)
    cnt = torch.tensor([s.counters()[k] for k in ("n_events", "n_pismc", "n_overflow")], dtype=torch.int64, device=acc.device)
    if world > 1:
        dist.all_reduce(cnt, group=group)
    out.n_events, out.n_pismc = int(cnt[0]), int(cnt[1])
    if int(cnt[2]):
        raise RuntimeError("a history outgrew the type-store capacity")
    s.close()
    return out


def shard_pairs(n_pairs: int, rank: int, world: int):
    """Indices of the locus pairs rank `rank` owns: round robin, so that neighbouring pairs (similar locus lengths,
    similar cost) spread over the ranks."""
    return list(range(int(rank), int(n_pairs), int(world)))


def gather_pair_results(mine, n_pairs: int, group=None):
    """mine: list of (pair index, result) of this rank -> list of all n_pairs results in pair order on every rank.
    One collective (all_gather_object: the results are a few small arrays per pair)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    parts = [mine]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine, group=group)
    out = [None] * n_pairs
    for part in parts:
        for k, r in part:
            assert out[k] is None, "pair %d computed twice" % k
            out[k] = r
    assert all(r is not None for r in out)
    return out


def run_sampler_batch_distributed(problems, num_iterations: int, seed: int = 0, quadrature_points: int = 16,
                                  device: Optional[int] = None, group=None, compute=None):
    """The double loop of main.cpp:242-311 over all ranks: with at least as many locus pairs as GPUs the pairs are the
    parallel axis -- rank r runs pairs r, r+R, ... through coalgpu_run_sampler_batch on its own GPU with no data-path
    communication, and the per-pair results are gathered once at the end. Every rank returns the full list, identical
    to single-GPU results (a pair's histories depend on (seed, index) only). `compute` replaces the GPU call in the
    CPU tests of the sharding/gather logic."""
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    idx = shard_pairs(len(problems), rank, world)
    if compute is None:
        from .sampler import ImportanceSampler
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(device)
        compute = lambda ps: ImportanceSampler().run_sampler_batch(ps, num_iterations, seed=seed, quadrature_points=quadrature_points, device=device)
    res = compute([problems[k] for k in idx]) if idx else []
    return gather_pair_results(list(zip(idx, res)), len(problems), group)
