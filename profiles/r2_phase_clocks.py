"""Phase clocks of the lockstep kernel (cycles between the CTA barriers as seen by thread 0: serial + cooperative sections S, grid
phase G, one-locus queries Q1, two-locus queries Q2), per lockstep step. Needs a library built with -DCOAL_LS_TIMING:
    COALGPU_LIB=variants/timing.so COALGPU_DEFS=-DCOAL_LS_TIMING python -m coalescenator_b200.build
    COALGPU_LIB=variants/timing.so python profiles/r2_phase_clocks.py 16 2 11 20      # bench pair indices
Results of the round: profiles/r2_kernel_iterations.md."""
import sys, ctypes as C
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch, bench
import coalescenator_b200 as cb
probs = bench.bench_problems()
idx = [int(a) for a in sys.argv[1:]] or [16, 2, 11]
for i in idx:
    p = probs[i]
    s = cb.DeviceSampler(p, device=0)
    n = 98304
    b = s.alloc(n)
    s.sample(b, 42, 0, 4096); torch.cuda.synchronize(); s.tune()
    out = (C.c_uint64 * 5)()
    s.L.coalgpu_phase_clocks.argtypes = [C.c_void_p, C.c_void_p]
    s.L.coalgpu_phase_clocks(s._h, out); base = list(out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.sample(b, 42, 100000, n); e1.record(); torch.cuda.synchronize()
    s.L.coalgpu_phase_clocks(s._h, out); d = [a - c for a, c in zip(list(out), base)]
    c = s.counters()
    tot = sum(d[:4])
    print("pair", i, "LA,LB", p.LA, p.LB, "ms %.1f" % e0.elapsed_time(e1), "hist/s %.0f" % (n / e0.elapsed_time(e1) * 1e3), "caps", c["store_first_pass"], c["lanes_per_history"],
          "steps/CTA-sum", d[4], "cycles/step: S %.0f G %.0f Q1 %.0f Q2 %.0f" % tuple(x / max(d[4], 1) for x in d[:4]),
          "share S %.2f G %.2f Q1 %.2f Q2 %.2f" % tuple(x / tot for x in d[:4]))
