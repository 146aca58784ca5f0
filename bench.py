#!/usr/bin/env python
"""Benchmark of the importance-sampling likelihood hot path (BASELINE.json metric:
importance-sampled histories/sec). Contract: see the task description / DESIGN.md §6.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the hot path over one batch: `--per-step` histories (default 98,304 per locus pair, the
1e5 importance samples per pair of BASELINE config 2) spread evenly over the bench locus pairs of configuration cfg2 (5 time points x 50 sequences x 10 kb, --loci-size 100),
every pair its own kernel launch + log-sum-exp reduction. Inputs (problem + HMM tables) are resident
in HBM for `value`; `e2e` goes through the drop-in entry point coalgpu_run_sampler with host buffers.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = "cfg2: 5 time points x 50 sequences x 10 kb synthetic alignment (seed 1), loci-size 100, G=1"
BENCH_PAIRS = (0, 40, 100, 170, 250, 320)   # indices into the eligible locus pairs of cfg2 (LA+LB from 3 to 11)


def bench_problems():
    from coalescenator_b200.problem import make_config, sub_problem
    aln, loci, pairs = make_config(2)
    return [sub_problem(aln, loci, *pairs[k]) for k in BENCH_PAIRS]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


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


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation (oracle/_ref/coalref, the unmodified sources
    compiled with header shims) on the box's host cores, same workload, bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle_py as O
    probs = bench_problems()
    cores = os.cpu_count() or 1
    kind = "reference" if O.ref_binary() else "port"
    per_pair = max(8 * cores, args.ref_per_pair)      # >= 8 histories per host thread, so table set-up per thread does not dominate
    def step():
        n = 0
        for p in probs:
            if kind == "reference":
                O.run_reference(p, per_pair, 42, threads=cores)
            else:
                O.run(p, per_pair, 42, mode=O.MODE_STRICT)
            n += per_pair
        return n
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    total = sum(step() for _ in range(args.steps))
    dt = time.perf_counter() - t0
    val = total / dt
    line = {"impl": "reference", "metric": "importance-sampled histories/sec", "value": val, "unit": "histories/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs": list(BENCH_PAIRS), "histories_per_step": per_pair * len(probs)},
            "cpu_baseline": {"value": val, "unit": "histories/s", "cores": cores if kind == "reference" else 1, "kind": kind,
                             "sample": "%d histories per locus pair x %d pairs per step, runSampler with %d threads" % (per_pair, len(probs), cores)},
            "e2e": {"value": val, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--per-step", type=int, default=589824,
                    help="histories per step PER GPU (weak scaling): 98,304 per bench locus pair, i.e. the 1e5 samples per pair of BASELINE config 2")
    ap.add_argument("--ref-per-pair", type=int, default=64)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import coalescenator_b200 as cb
    from coalescenator_b200.distributed import combine_partials, merge_pairs, empty_partials
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

    probs = bench_problems()
    npairs = len(probs)
    per_pair = args.per_step // npairs
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

    # dominant kernel: history_kernel, timed live with CUDA events on the launching stream
    k_ms = [a.elapsed_time(b) for a, b in hist_ms]
    k_avg_ms = float(np.mean(k_ms))
    cnt = [s.counters() for s in samplers]
    ev_total = sum(c["n_events"] for c in cnt)
    pi_total = sum(c["n_pismc"] for c in cnt)
    n_drawn = per_pair * npairs * (args.warmup + args.steps)
    assert sum(c["n_overflow"] for c in cnt) == 0

    # e2e: the drop-in call (host problem in, host result out; tables built + uploaded inside the call)
    e2e_n = per_pair
    t0 = time.perf_counter()
    h2d = d2h = 0
    e2e_dev = 0.0
    for p in probs:
        r = cb.ImportanceSampler().run_sampler(p, e2e_n, seed=4242 + rank, device=local)
        e2e_dev += r.seconds_device
        d2h += (3 + p.T) * p.G * 2 * 8
        h2d += sum(len(tp) for tp in p.types) * 16 + (p.T * 8 + p.G * p.T * 6) * 8
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    te = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = e2e_n * npairs * world / float(te.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_kind = peaks()
    T = probs[0].T
    # HBM roofline (the contract's bound): algorithmic bytes of one launch = what a history must leave in HBM,
    # (G + 1 + T) doubles + its event count; problem and tables are shared by all histories.
    bytes_per_hist = (1 + 1 + T) * 8 + 4
    achieved = per_pair * bytes_per_hist / (k_avg_ms * 1e-3) / 1e9
    # What actually binds is the fp64 pipe / instruction issue. Algorithmic FLOPs per launch (SURVEY.md §8d,
    # "minimum after hoisting"): P_C (2 Q^2 + 8 nH Q) + P_AB 4 nH Q with the counters of this very run.
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
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, scaled to this launch size
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["histories_in_launch"] * per_pair
        traffic_note = "ncu capture of a %d-history launch (%s), scaled to %d histories" % (tj["histories_in_launch"], tj["capture"], per_pair)
    roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
            "traffic": traffic, "traffic_note": traffic_note, "algorithmic_bytes": per_pair * bytes_per_hist, "peak_kind": pk_kind, "kernel": "history_kernel", "kernel_ms": k_avg_ms,
            "note": "path is fp64-ALU / issue bound, not HBM bound (SURVEY.md §8d): see alu",
            "alu": {"bound": "fp64", "achieved": fp64_achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_achieved / fp64_peak,
                    "peak_kind": "measured in this run (coalgpu_fp64_peak, dependent FMA chains)",
                    "flop_model": "P_C*(2Q^2+8*nH*Q) + P_AB*4*nH*Q (SURVEY.md 8d)",
                    "events_per_history": ev_total / n_drawn, "pismc_per_history": pi_total / n_drawn,
                    "two_locus_per_history": p_c / n_drawn, "classes_per_two_locus": nh,
                    "pismc_per_s": pi_total / n_drawn * per_pair / (k_avg_ms * 1e-3)}}
    line = {"metric": "importance-sampled histories/sec", "value": value, "unit": "histories/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs": list(BENCH_PAIRS), "histories_per_step": per_pair * npairs * world,
                       "l2": "per-history state is on chip; every step writes fresh outputs (inputs are L2-resident tables by design)",
                       "parallelism": "dp%d over histories" % world,
                       "capacity": "two-pass type store tuned after warm-up step 1: (first-pass store, full store, lanes per history, peak types) per pair = %s"
                                   % ([(c["store_first_pass"], c["store_full"], c["lanes_per_history"], c["peak_types"]) for c in caps] if caps else "untuned")},
            "clocks": clk.summary(), "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": "histories/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "histories": e2e_n * npairs * world, "wall_s_rank0": e2e_dt, "device_s_rank0": e2e_dev},
            "roofline": roof}

    if not args.no_cpu_baseline:
        from oracle import oracle_py as O
        cores = os.cpu_count() or 1
        kind = "reference" if O.ref_binary() else "port"
        n_cpu, t_cpu = 0, 0.0
        per = 8 * cores                                  # 8 histories per host thread
        for p in probs:
            t0 = time.perf_counter()
            if kind == "reference":
                O.run_reference(p, per, 42, threads=cores)
            else:
                O.run(p, per, 42, mode=O.MODE_STRICT)
            t_cpu += time.perf_counter() - t0
            n_cpu += per
            if t_cpu > args.cpu_seconds:
                break
        line["cpu_baseline"] = {"value": n_cpu / t_cpu, "unit": "histories/s", "cores": cores if kind == "reference" else 1, "kind": kind,
                                "sample": "%d histories per locus pair over %d of the %d bench pairs" % (per, n_cpu // per, npairs)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
