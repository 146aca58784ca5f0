#!/usr/bin/env python
"""Benchmark of the importance-sampling likelihood hot path (BASELINE.json metric:
importance-sampled histories/sec). Contract: see the task description / DESIGN.md §6.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfg2|cfg3|cfg4]

A step = one pass of the hot path over one batch: every bench locus pair of the configuration draws its
histories (one history-kernel launch or two per pair + the log-sum-exp reduction). Inputs (problem + HMM
tables) are resident in HBM for `value`; `e2e` goes through the drop-in entry point coalgpu_run_sampler
with host buffers. Configurations (BASELINE.json `configs`, synthetic inputs of SURVEY.md §8d):
  cfg2 (default, the metric's configuration): 5 time points x 50 sequences x 10 kb, G = 1, 98,304 histories per pair
       (the 1e5 samples per pair of config 2) over a SEEDED RANDOM SAMPLE of 24 of the eligible locus pairs;
  cfg3: the cfg2 data with the 32 x 1 x 32 (mu, rho, lambda) target grid (G = 1024), 10,000 histories per pair, 32 pairs;
  cfg4: 10 time points x 200 sequences x 100 kb, 8 pairs x 131,072 histories = 1,048,576 (~1e6) histories per step.
With --gpus N > 1 (torchrun) every rank runs the same pairs on its own index ranges (weak scaling) and the ranks
exchange their (max, sum) partials with one NCCL collective per step; after the timed region rank 0..N-1 also run the
sample-sharded path on one pair and compare its grids with the single-GPU call (`multi_gpu_check`).
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PAIR_SEED = 20261020
CONFIGS = {
    # name: (make_config id, pairs in the sample, histories per pair per step, grid)
    "cfg2": dict(cfg=2, n_pairs=24, per_pair=98304, grid=None,
                 workload="cfg2: 5 time points x 50 sequences x 10 kb synthetic alignment (seed 1), loci-size 100, G=1; "
                          "bench pairs = random.Random(%d).sample of 24 of the eligible locus pairs, 98,304 histories per pair per step" % PAIR_SEED),
    "cfg3": dict(cfg=2, n_pairs=32, per_pair=10000, grid="cfg3",
                 workload="cfg3: cfg2 data, 32 x 1 x 32 (mu, rho, lambda) target grid (G=1024); "
                          "bench pairs = random.Random(%d).sample of 32 of the eligible locus pairs, 10,000 histories per pair per step" % PAIR_SEED),
    "cfg4": dict(cfg=4, n_pairs=8, per_pair=131072, grid=None,
                 workload="cfg4: 10 time points x 200 sequences x 100 kb synthetic alignment (seed 1), loci-size 100, G=1; "
                          "bench pairs = random.Random(%d).sample of 8 of the eligible locus pairs, 131,072 histories per pair per step (1,048,576 per step)" % PAIR_SEED),
}
_cache = {}


def bench_pairs(config="cfg2"):
    """Indices (into the eligible locus pairs of the configuration, main.cpp:242-252) of the bench pairs: a seeded
    random sample, so the mix of locus lengths / haplotype diversity is the data set's, not a hand-picked one."""
    c = CONFIGS[config]
    if ("pairs", config) not in _cache:
        from coalescenator_b200.problem import make_config
        aln, loci, pairs = make_config(c["cfg"])
        eligible = [k for k, (i, j, d) in enumerate(pairs) if (loci[i][3] - loci[i][2]) + (loci[j][3] - loci[j][2]) <= 64]
        _cache[("pairs", config)] = (aln, loci, pairs, sorted(random.Random(PAIR_SEED).sample(eligible, c["n_pairs"])))
    return _cache[("pairs", config)]


def bench_problems(config="cfg2"):
    from coalescenator_b200.problem import sub_problem, grid_cfg3
    c = CONFIGS[config]
    aln, loci, pairs, idx = bench_pairs(config)
    kw = {}
    if c["grid"] == "cfg3":
        mu, rho, lam = grid_cfg3()
        kw = dict(target_mu=mu, target_rho=rho, target_lambda=lam)
    return [sub_problem(aln, loci, *pairs[k], **kw) for k in idx]


BENCH_PAIRS = None   # kept for callers of the round-1 module attribute: see bench_pairs()


def config_block(config):
    """The `config` object of the JSON line: identical in both arms (the driver compares them)."""
    return {"workload": CONFIGS[config]["workload"], "pairs": list(bench_pairs(config)[3])}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.stop = threading.Event()
        self.thread = None

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


def time_reference(probs, per_thread, cores, kind):
    """The reference's CPU sampler on `probs`, `per_thread` histories per host thread and pair. Returns
    (histories, seconds inside runSampler, wall seconds): `seconds` is the driver's own clock around
    ImportanceSampler::runSampler only (oracle/ref_driver.cpp), i.e. without process start and input parsing."""
    from oracle import oracle_py as O
    n = in_sampler = 0
    t0 = time.perf_counter()
    for p in probs:
        if kind == "reference":
            r = O.run_reference(p, per_thread * cores, 42, threads=cores)
            in_sampler += float(r["seconds"])
        else:
            t1 = time.perf_counter()
            O.run(p, per_thread, 42, mode=O.MODE_STRICT)
            in_sampler += time.perf_counter() - t1
        n += per_thread * (cores if kind == "reference" else 1)
    return n, in_sampler, time.perf_counter() - t0


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation (oracle/_ref/coalref, the unmodified sources
    compiled with header shims) on the box's host cores, same workload, bounded sample per step. The rate is
    measured at TWO sample sizes per host thread (set-up -- every thread rebuilds its HMM tables,
    ImportanceSampler.cpp:89-91 -- must be amortised for the figure to mean anything); `value` is the better
    of the two rates, timed by the driver's own clock around runSampler."""
    if rank != 0:
        return
    from oracle import oracle_py as O
    probs = bench_problems(args.config)
    cores = os.cpu_count() or 1
    kind = "reference" if O.ref_binary() else "port"
    small, large = args.ref_per_thread_small, args.ref_per_thread
    # bounded sample: a subset of the bench pairs per step (all of them share the data set; the subset is the first
    # `ref_pairs` of the seeded sample)
    sub = probs[:max(1, min(len(probs), args.ref_pairs))]
    for _ in range(args.warmup):
        time_reference(sub[:1], small, cores, kind)
    t0 = time.perf_counter()
    tot_n = tot_s = 0.0
    for _ in range(args.steps):
        n, s, _w = time_reference(sub, large, cores, kind)
        tot_n += n; tot_s += s
    dt = time.perf_counter() - t0
    val_large = tot_n / tot_s
    n_s, s_s, _w = time_reference(sub[:2], small, cores, kind)
    val_small = n_s / s_s
    val = max(val_large, val_small)       # the reference's better rate: conservative for every speed-up quoted against it
    line = {"impl": "reference", "metric": "importance-sampled histories/sec", "value": val, "unit": "histories/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(args.config),
            "cpu_baseline": {"value": val, "unit": "histories/s", "cores": cores if kind == "reference" else 1, "kind": kind,
                             "sample": "%d histories per host thread and locus pair over the first %d bench pairs per step, runSampler with %d threads; "
                                       "timed by the driver's clock around runSampler (wall clock incl. process start: %.1f histories/s)"
                                       % (large, len(sub), cores, tot_n / dt),
                             "amortisation": {"histories_per_thread": [small, large], "histories_per_s": [val_small, val_large],
                                              "ratio_small_over_large": val_small / val_large,
                                              "note": "`value` is the HIGHER of the two rates (set-up is amortised at both sizes: the larger sample is not faster)"}},
            "e2e": {"value": val, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "details": {"histories_per_step": int(tot_n / max(args.steps, 1))}}
    print(json.dumps(line))


def problem_bytes(p):
    """Bytes one drop-in call moves: host -> device = everything coalgpu_run_sampler uploads for the problem (the arena of
    coal_api.cu:finish_device -- the HMM table rows it builds on the host inside the call, per-epoch and per-grid-point
    constants, log-factorial / reciprocal tables, the merged type lists), device -> host = the (max, sum) partials and the
    counters it reads back per chunk."""
    n_max = sum(c * (2 if g == 2 else 1) for tp in p.types for (g, _w, c) in tp)
    n_sets = len({v for v in p.viral})
    dev_row = (2 * (p.LA + 2) * (p.LB + 2) + p.LA + p.LB + 2 + 1) & ~1
    n_types = sum(len({(g, w) for (g, w, _c) in tp}) for tp in p.types)
    h2d = (n_sets * n_max * dev_row + p.T + 8 * p.T + 6 * p.T * p.G + 2 * (n_max + 2)) * 8 + (p.T + p.T + 1) * 4 + n_types * 12
    chunks = 1     # one chunk at the bench sizes (coal_api.cu:sample_range; a pilot chunk is added only for long runs)
    d2h = chunks * (3 + p.T) * p.G * 2 * 8 + 16 * 8
    return h2d, d2h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--per-pair", type=int, default=None, help="histories per bench pair per step PER GPU (default: the configuration's)")
    ap.add_argument("--ref-per-thread", type=int, default=32, help="reference arm: histories per host thread and pair (large sample)")
    ap.add_argument("--ref-per-thread-small", type=int, default=8)
    ap.add_argument("--ref-pairs", type=int, default=3, help="reference arm / cpu_baseline: bench pairs per step")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 5 if args.config == "cfg2" else 3

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import coalescenator_b200 as cb
    from coalescenator_b200.distributed import combine_partials, run_sampler_distributed
    from coalescenator_b200.sampler import launch_count

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: what NCCL / torch print while the communicator comes up
        # ("NCCL version ..." under NCCL_DEBUG=VERSION) is sent to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    cfg = CONFIGS[args.config]
    probs = bench_problems(args.config)
    npairs = len(probs)
    per_pair = args.per_pair or cfg["per_pair"]
    samplers = [cb.DeviceSampler(p, device=local) for p in probs]
    bufs = [s.alloc(per_pair) for s in samplers]
    stream = torch.cuda.current_stream(local)
    hist_ms = []

    def step(k, timed=False):
        """One pass: every pair draws per_pair histories on this rank (disjoint Philox index ranges per rank and
        step), reduces them to (max, sum) partials, and the ranks exchange the partials with one collective."""
        outs = []
        for i, (s, b) in enumerate(zip(samplers, bufs)):
            first = (k * world + rank) * per_pair
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            s.sample(b, 42, first, per_pair)
            if timed:
                e1.record(stream)
                hist_ms.append((e0, e1))
            outs.append(s.partials(b, per_pair))
        allp = torch.stack(outs)
        return combine_partials(allp) if world > 1 else allp

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    caps = None
    for k in range(args.warmup):
        step(k)
        if k == 0:
            # size the reduced type store / lanes per history from what the first warm-up histories needed
            # (coalgpu_tune; results never depend on it). coalgpu_run_sampler (the e2e leg) does the same after a pilot chunk.
            sync()
            caps = [s.tune() for s in samplers]
    sync()
    l0 = launch_count()
    with ClockSampler(local) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        e0.record(stream)
        for k in range(args.steps):
            last = step(args.warmup + k, timed=True)
        e1.record(stream)
        sync()
        ms = e0.elapsed_time(e1)
    launches = launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    total_hist = per_pair * npairs * world * args.steps
    value = total_hist / (ms * 1e-3)

    # dominant kernel: history_kernel (first + second pass of a pair), timed live with CUDA events on the launching stream
    k_ms = [a.elapsed_time(b) for a, b in hist_ms]
    k_avg_ms = float(np.mean(k_ms))
    k_share = float(np.sum(k_ms)) / ms if world == 1 else None
    cnt = [s.counters() for s in samplers]
    ev_total = sum(c["n_events"] for c in cnt)
    pi_total = sum(c["n_pismc"] for c in cnt)
    n_drawn = per_pair * npairs * (args.warmup + args.steps)
    assert sum(c["n_overflow"] for c in cnt) == 0

    # e2e: the drop-in call (host problem in, host result out; tables built + uploaded inside the call), one pass cold
    # (first call of these problems in the process: table cache empty) and one pass warm
    e2e = None
    if not args.no_e2e:
        h2d = d2h = 0
        for p in probs:
            a, b = problem_bytes(p)
            h2d += a; d2h += b
        passes = []
        from coalescenator_b200.sampler import transfer_bytes
        for rep in range(2):
            tb0 = transfer_bytes()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            e2e_dev = 0.0
            for p in probs:
                r = cb.ImportanceSampler().run_sampler(p, per_pair, seed=4242 + rank, device=local)
                e2e_dev += r.seconds_device
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            te = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            tb1 = transfer_bytes()
            passes.append((per_pair * npairs * world / float(te.item()), dt, e2e_dev, tb1[0] - tb0[0], tb1[1] - tb0[1]))
        # For information: the same pairs through ONE coalgpu_run_sampler_batch call (what replaces the reference's whole pair loop,
        # main.cpp:242-311: pairs on streams, host preparation overlapped; the drain of one pair's launch is covered by the next
        # pair's). Not the headline: `value` above is the per-pair drop-in call, the smaller change for a maintainer.
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        cb.ImportanceSampler().run_sampler_batch(probs, per_pair, seed=4242 + rank, device=local)
        torch.cuda.synchronize()
        tb = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        batch_value = per_pair * npairs * world / float(tb.item())
        # bytes per step: MEASURED by the library (coalgpu_transfer_bytes: every cudaMemcpy of the drop-in call counts its bytes),
        # this rank, warm pass; `bytes_model` is the closed-form count of problem_bytes() for comparison
        e2e = {"value": passes[1][0], "unit": "histories/s", "h2d_bytes_per_step": int(passes[1][3]), "d2h_bytes_per_step": int(passes[1][4]),
               "bytes": "measured by the library (coalgpu_transfer_bytes), rank 0, warm pass", "bytes_model": [int(h2d), int(d2h)],
               "batch_call_value": batch_value,
               "histories": per_pair * npairs * world, "cold_value": passes[0][0], "warm_value": passes[1][0],
               "wall_s_rank0": passes[1][1], "device_s_rank0": passes[1][2]}

    # N > 1: the sample-sharded path on one pair against the single-GPU call of the same (seed, n) -- multi-GPU
    # CORRECTNESS where the driver's scaling run can see it (VERDICT r1 task 1a)
    mg = None
    if world > 1:
        n_chk = 16384
        chk = run_sampler_distributed(probs[0], n_chk, seed=77, device=local)
        one = cb.ImportanceSampler().run_sampler(probs[0], n_chk, seed=77, device=local)
        def rel(a, b):
            return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
        d = max(rel(chk.likelihood, one.likelihood), rel(chk.likelihoodSq, one.likelihoodSq), rel(chk.tmrca, one.tmrca),
                rel(chk.lineages_remaining, one.lineages_remaining))
        ok_local = d < 1e-10 and chk.n_events == one.n_events
        flag = torch.tensor([0 if ok_local else 1], dtype=torch.int64, device=dev)
        dmax = torch.tensor([d], dtype=torch.float64, device=dev)
        dist.all_reduce(flag)
        dist.all_reduce(dmax, op=dist.ReduceOp.MAX)
        mg = {"max_rel_diff": float(dmax.item()), "ok": int(flag.item()) == 0, "histories": n_chk, "ranks": world,
              "what": "run_sampler_distributed (rank r draws shard_range(n, r, R), NCCL all_gather of the (max, sum) pairs) vs the "
                      "single-GPU drop-in call with the same (seed, n), on every rank; likelihood, likelihoodSq, tmrca, lineages, event count"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        if mg is not None and not mg["ok"]:
            sys.exit(3)
        return

    pk, pk_kind = peaks()
    T = probs[0].T
    G = probs[0].G
    # HBM roofline (the contract's bound): algorithmic bytes of one launch = what a history must leave in HBM,
    # (G + 1 + T) doubles + its event count; problem and tables are shared by all histories.
    bytes_per_hist = (G + 1 + T) * 8 + 4
    achieved = per_pair * bytes_per_hist / (k_avg_ms * 1e-3) / 1e9
    # What actually binds is instruction issue (fp64 + integer); algorithmic FLOPs per launch in the model of SURVEY.md §8d
    # ("minimum after hoisting": P_C (2 Q^2 + 8 nH Q) + P_AB 4 nH Q) with the counters of this very run, so that the figure
    # stays comparable with round 1. The kernel itself now executes FEWER flops than that model (the 16-interval forward
    # pass is a bilinear form of ~2 nH^2 multiply-adds), so `frac` is work-equivalent throughput, not pipe use.
    Q = 16
    p_c = sum(c["n_two_locus"] for c in cnt)
    n_cls = sum(c["n_classes"] for c in cnt)
    nh = n_cls / max(p_c, 1)
    p_ab = pi_total - p_c
    flop_total = p_c * 2 * Q * Q + 8 * Q * n_cls + p_ab * 4 * nh * Q
    flop_per_launch = flop_total / (n_drawn / per_pair)
    from coalescenator_b200.sampler import fp64_peak_tflops
    fp64_peak = fp64_peak_tflops(local)
    fp64_achieved = flop_per_launch / (k_avg_ms * 1e-3) / 1e12
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, scaled to this launch size
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["histories_in_launch"] * per_pair
        traffic_note = "ncu capture of a %d-history launch (%s), scaled to %d histories" % (tj["histories_in_launch"], tj["capture"], per_pair)
    roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
            "traffic": traffic, "traffic_note": traffic_note, "algorithmic_bytes": per_pair * bytes_per_hist, "peak_kind": pk_kind, "kernel": "history_kernel", "kernel_ms": k_avg_ms,
            "kernel_share_of_step": k_share,
            "note": "path is instruction-issue / fp64 bound, not HBM bound (SURVEY.md §8d): see alu",
            "alu": {"bound": "fp64", "achieved": fp64_achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_achieved / fp64_peak,
                    "peak_kind": "measured in this run: coalgpu_fp64_peak = 8 independent dependent-FMA chains per thread, 256 threads x 8 CTAs per SM, "
                                 "4096 x 16 x 8 DFMA per thread, best of 3 timed launches after one warm-up, CUDA events, clocks as in `clocks` "
                                 "(MEASURED_PEAKS.json has no fp64 entry)",
                    "flop_model": "P_C*(2Q^2+8*nH*Q) + P_AB*4*nH*Q (SURVEY.md 8d; work-equivalent: the kernel evaluates the forward pass as a bilinear form)",
                    "events_per_history": ev_total / n_drawn, "pismc_per_history": pi_total / n_drawn,
                    "two_locus_per_history": p_c / n_drawn, "classes_per_two_locus": nh,
                    "pismc_per_s": pi_total / n_drawn * per_pair / (k_avg_ms * 1e-3)}}
    line = {"metric": "importance-sampled histories/sec", "value": value, "unit": "histories/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(args.config),
            "details": {"histories_per_step": per_pair * npairs * world,
                        "l2": "per-history state is on chip; every step writes fresh outputs (inputs are L2-resident tables by design)",
                        "parallelism": "dp%d over histories" % world,
                        "capacity": "two-pass type store tuned after warm-up step 1: (first-pass store, full store, lanes per history, peak types) per pair = %s"
                                    % ([(c["store_first_pass"], c["store_full"], c["lanes_per_history"], c["peak_types"]) for c in caps] if caps else "untuned")},
            "clocks": clk.summary(), "gpu_launches": int(launches),
            "roofline": roof}
    if e2e is not None:
        line["e2e"] = e2e
    if mg is not None:
        line["multi_gpu_check"] = mg

    if not args.no_cpu_baseline:
        from oracle import oracle_py as O
        cores = os.cpu_count() or 1
        kind = "reference" if O.ref_binary() else "port"
        n_cpu, t_cpu, used = 0, 0.0, 0
        for p in probs[:max(1, args.ref_pairs)]:
            n, s, _w = time_reference([p], 8, cores, kind)
            n_cpu += n; t_cpu += s; used += 1
            if t_cpu > args.cpu_seconds:
                break
        line["cpu_baseline"] = {"value": n_cpu / t_cpu, "unit": "histories/s", "cores": cores if kind == "reference" else 1, "kind": kind,
                                "sample": "8 histories per host thread over the first %d of the %d bench pairs (driver clock around runSampler); "
                                          "`bench.py --impl reference` times the larger sample" % (used, npairs)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if mg is not None and not mg["ok"]:
        sys.exit(3)


if __name__ == "__main__":
    main()
