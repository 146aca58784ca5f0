"""Host-side mirror of the reference's sampler interface, bound to libcoalgpu.so over its C ABI.

Reference interface mirrored here (same names, argument meaning and result fields):
  ImportanceSampler::runSampler(Sample*, ModelParameters, num_iterations, num_threads,
                                quadrature_points, seed) -> SamplerOutput<double>
      /root/reference/src/ImportanceSampler.hpp:53, ImportanceSampler.cpp:110-145
  SamplerOutput<double>::likelihood() / likelihoodSq() / tmrca() / lineages_remaining()
      /root/reference/src/SamplerOutput.hpp:44-52, read at main.cpp:285-306

There is no CPU path: if the CUDA library is missing or no B200 is visible these calls raise.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
from typing import List, Optional

import numpy as np

from .problem import Problem

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COALGPU_LIB") or os.path.join(HERE, "libcoalgpu.so")   # COALGPU_LIB: tuning variants only

EVENT_DTYPE = np.dtype([("chosen_word", "<u8"), ("aux_word", "<u8"), ("chosen_gamma", "u1"), ("kind", "u1"),
                        ("epoch", "u1"), ("site", "u1"), ("n_before", "<u4")])


class CoalGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"coalgpu error {code}: {msg}")
        self.code = code


class _CProblem(C.Structure):
    _fields_ = [("T", C.c_int32), ("LA", C.c_int32), ("LB", C.c_int32),
                ("times", C.POINTER(C.c_double)), ("viral", C.POINTER(C.c_double)),
                ("tau", C.c_double), ("driving_mu", C.c_double), ("driving_lambda", C.c_double), ("driving_rho", C.c_double),
                ("n_mu", C.c_int32), ("n_rho", C.c_int32), ("n_lambda", C.c_int32),
                ("mu", C.POINTER(C.c_double)), ("rho", C.POINTER(C.c_double)), ("lam", C.POINTER(C.c_double)),
                ("type_offsets", C.POINTER(C.c_int32)), ("type_gamma", C.POINTER(C.c_uint8)),
                ("type_words", C.POINTER(C.c_uint64)), ("type_counts", C.POINTER(C.c_uint32))]


class _CResult(C.Structure):
    _fields_ = [("likelihood", C.POINTER(C.c_double)), ("likelihood_sq", C.POINTER(C.c_double)),
                ("tmrca", C.POINTER(C.c_double)), ("lineages", C.POINTER(C.c_double)), ("ess", C.POINTER(C.c_double)),
                ("n_histories", C.c_uint64), ("n_events", C.c_uint64), ("n_pismc", C.c_uint64),
                ("seconds_device", C.c_double)]


_lib = None

EXPORTS = ["coalgpu_run_sampler", "coalgpu_run_sampler_batch", "coalgpu_create", "coalgpu_destroy", "coalgpu_sample", "coalgpu_partials",
           "coalgpu_finalize", "coalgpu_counters", "coalgpu_tables", "coalgpu_pismc_classes",
           "coalgpu_launch_count", "coalgpu_transfer_bytes", "coalgpu_last_error", "coalgpu_version", "coalgpu_work_counters", "coalgpu_fp64_peak",
           "coalgpu_capacity_stats", "coalgpu_tune", "coalgpu_set_emission_mode",
           "coalgpu_comm_init", "coalgpu_comm_finalize", "coalgpu_comm_size", "coalgpu_run_sampler_comm",
           "coalgpu_run_sampler_multi", "coalgpu_multi_shutdown", "coalgpu_nccl_calls", "coalgpu_score_events"]


def load_library():
    """dlopen libcoalgpu.so; fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.coalgpu_last_error.restype = C.c_char_p
        L.coalgpu_version.restype = C.c_char_p
        L.coalgpu_launch_count.restype = C.c_uint64
        L.coalgpu_sample.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.coalgpu_partials.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.coalgpu_create.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]
        L.coalgpu_destroy.argtypes = [C.c_void_p]
        L.coalgpu_destroy.restype = None
        L.coalgpu_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.coalgpu_run_sampler.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p]
        L.coalgpu_run_sampler_batch.argtypes = [C.c_int32, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p]
        L.coalgpu_finalize.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.coalgpu_tables.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 5
        L.coalgpu_pismc_classes.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 8
        L.coalgpu_work_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.coalgpu_capacity_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)] + [C.POINTER(C.c_uint32)] * 4
        L.coalgpu_tune.argtypes = [C.c_void_p]
        L.coalgpu_set_emission_mode.argtypes = [C.c_int]
        L.coalgpu_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.coalgpu_comm_init.argtypes = [C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]
        L.coalgpu_comm_finalize.argtypes = [C.c_void_p]
        L.coalgpu_comm_finalize.restype = None
        L.coalgpu_comm_size.argtypes = [C.c_void_p]
        L.coalgpu_run_sampler_comm.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p]
        L.coalgpu_run_sampler_multi.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int32, C.c_void_p, C.c_void_p]
        L.coalgpu_multi_shutdown.restype = None
        L.coalgpu_nccl_calls.restype = C.c_uint64
        L.coalgpu_score_events.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 7 + [C.c_int32] + [C.c_void_p] * 3
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise CoalGpuError(rc, load_library().coalgpu_last_error().decode())


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class _PackedProblem:
    """numpy arrays behind a coalgpu_problem struct (kept alive for the duration of the calls)."""

    def __init__(self, p: Problem):
        self.times = np.asarray(p.times, dtype=np.float64)
        self.viral = np.asarray(p.viral, dtype=np.float64)
        self.mu = np.asarray(p.target_mu, dtype=np.float64)
        self.rho = np.asarray(p.target_rho, dtype=np.float64)
        self.lam = np.asarray(p.target_lambda, dtype=np.float64)
        offs, g, w, c = [0], [], [], []
        for tp in p.types:
            for (gg, ww, cc) in tp:
                g.append(gg); w.append(ww); c.append(cc)
            offs.append(len(g))
        self.offs = np.asarray(offs, dtype=np.int32)
        self.g = np.asarray(g, dtype=np.uint8)
        self.w = np.asarray(w, dtype=np.uint64)
        self.c = np.asarray(c, dtype=np.uint32)
        self.struct = _CProblem(p.T, p.LA, p.LB, _dp(self.times), _dp(self.viral), p.tau, p.driving_mu,
                                p.driving_lambda, p.driving_rho, len(self.mu), len(self.rho), len(self.lam),
                                _dp(self.mu), _dp(self.rho), _dp(self.lam),
                                self.offs.ctypes.data_as(C.POINTER(C.c_int32)), self.g.ctypes.data_as(C.POINTER(C.c_uint8)),
                                self.w.ctypes.data_as(C.POINTER(C.c_uint64)), self.c.ctypes.data_as(C.POINTER(C.c_uint32)))


@dataclasses.dataclass
class SamplerOutput:
    """SamplerOutput<double> of the reference (SamplerOutput.hpp:40-69), log domain, arrays shaped
    (n_mu, n_rho, n_lambda); lineages_remaining has a leading axis over sampling times."""
    likelihood: np.ndarray
    likelihoodSq: np.ndarray
    tmrca: np.ndarray
    lineages_remaining: np.ndarray
    ess: np.ndarray
    n_histories: int = 0
    n_events: int = 0
    n_pismc: int = 0
    seconds_device: float = 0.0

    def standard_error(self):
        """Monte-Carlo standard error of `likelihood` (log domain, delta method) from the two sums the
        reference prints (main.cpp:240,306): var(log mean w) ~ (sum w^2 / (sum w)^2 - 1/N)."""
        rel = np.exp(self.likelihoodSq - 2 * self.likelihood) - 1.0 / max(self.n_histories, 1)
        return np.sqrt(np.maximum(rel, 0.0))


def set_emission_mode(mode: str) -> None:
    """'reference' (default: tables bit-identical to the reference's), 'stable' (closed form by quadrature, every
    entry accurate at any locus length; built by a kernel on the device) or 'stable-host' (the same quadrature on the
    host); applies to samplers created afterwards (coalgpu_set_emission_mode)."""
    _check(load_library().coalgpu_set_emission_mode({"reference": 0, "stable": 1, "stable-host": 2}[mode]))


def _result_arrays(p: Problem):
    G, T = p.G, p.T
    return (np.zeros(G), np.zeros(G), np.zeros(G), np.zeros(T * G), np.zeros(G))


def _to_output(p: Problem, arrs, res: Optional[_CResult] = None) -> SamplerOutput:
    shape = (len(p.target_mu), len(p.target_rho), len(p.target_lambda))
    like, likesq, tm, lin, ess = arrs
    out = SamplerOutput(like.reshape(shape), likesq.reshape(shape), tm.reshape(shape), lin.reshape((p.T,) + shape), ess.reshape(shape))
    if res is not None:
        out.n_histories, out.n_events, out.n_pismc = int(res.n_histories), int(res.n_events), int(res.n_pismc)
        out.seconds_device = float(res.seconds_device)
    return out


class ImportanceSampler:
    """Drop-in for the reference's ImportanceSampler (ImportanceSampler.hpp:42-53)."""

    def run_sampler(self, sample: Problem, num_iterations: int, num_threads: int = 1, quadrature_points: int = 16,
                    seed: int = 0, device: int = 0) -> SamplerOutput:
        """runSampler (ImportanceSampler.cpp:110-145) through the C ABI: host buffers in and out, blocking.
        `num_threads` is accepted for signature parity and ignored (the GPU draws all histories)."""
        L = load_library()
        pk = _PackedProblem(sample)
        arrs = _result_arrays(sample)
        res = _CResult(_dp(arrs[0]), _dp(arrs[1]), _dp(arrs[2]), _dp(arrs[3]), _dp(arrs[4]), 0, 0, 0, 0.0)
        _check(L.coalgpu_run_sampler(C.byref(pk.struct), int(num_iterations), int(quadrature_points), int(seed), int(device), C.byref(res)))
        return _to_output(sample, arrs, res)

    runSampler = run_sampler

    def run_sampler_batch(self, samples, num_iterations: int, quadrature_points: int = 16, seed: int = 0, device: int = 0):
        """runSampler for every locus pair of an analysis (the loop of main.cpp:242-311) in one call: the
        pairs run concurrently on the GPU. Returns one SamplerOutput per problem, identical to run_sampler's."""
        L = load_library()
        n = len(samples)
        packed = [_PackedProblem(p) for p in samples]
        probs = (_CProblem * n)(*[pk.struct for pk in packed])
        arrs = [_result_arrays(p) for p in samples]
        ress = (_CResult * n)(*[_CResult(_dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(a[4]), 0, 0, 0, 0.0) for a in arrs])
        _check(L.coalgpu_run_sampler_batch(n, probs, int(num_iterations), int(quadrature_points), int(seed), int(device), ress))
        return [_to_output(p, a, r) for p, a, r in zip(samples, arrs, ress)]


    def run_sampler_multi(self, sample: Problem, num_iterations: int, devices=None, quadrature_points: int = 16,
                          seed: int = 0) -> SamplerOutput:
        """runSampler with the importance samples of ONE locus pair sharded over several GPUs of this process
        (coalgpu_run_sampler_multi: one host thread per device, NCCL allreduce of the (max, sum) pairs inside the
        library). `devices`: list of CUDA device indices (default: all visible). Same histories, hence the same
        estimates up to merge rounding, as run_sampler(num_iterations, seed) on one device -- the thread fan-out and
        merge of ImportanceSampler.cpp:110-145 with GPUs in the place of threads."""
        L = load_library()
        if devices is None:
            import torch
            devices = list(range(torch.cuda.device_count()))
        dv = np.asarray(list(devices), dtype=np.int32)
        pk = _PackedProblem(sample)
        arrs = _result_arrays(sample)
        res = _CResult(_dp(arrs[0]), _dp(arrs[1]), _dp(arrs[2]), _dp(arrs[3]), _dp(arrs[4]), 0, 0, 0, 0.0)
        _check(L.coalgpu_run_sampler_multi(C.byref(pk.struct), int(num_iterations), int(quadrature_points), int(seed),
                                           len(dv), dv.ctypes.data_as(C.c_void_p), C.byref(res)))
        return _to_output(sample, arrs, res)

    def run_sampler_driving_grid(self, sample: Problem, num_iterations: int, quadrature_points: int = 16, seed: int = 0,
                                 device: int = 0) -> SamplerOutput:
        """The optional variant of SURVEY.md section 8d.3: ONE DRIVING POINT PER GRID POINT. The reference scores every
        (mu, rho, lambda) of the target grid from histories proposed under a single driving point, and the weights
        are heavy-tailed away from it (effective sample sizes of a few units on most of a grid even with 60,000
        histories). Here every grid point proposes its own histories: the grid points become the problems of one
        coalgpu_run_sampler_batch call (an extra batch axis), num_iterations histories each. Returns a SamplerOutput
        shaped like run_sampler's; entry (i, j, k) is what run_sampler returns for driving = target = that point."""
        import dataclasses
        probs = []
        for mu in sample.target_mu:
            for rho in sample.target_rho:
                for lam in sample.target_lambda:
                    probs.append(dataclasses.replace(sample, driving_mu=mu, driving_rho=rho, driving_lambda=lam,
                                                     target_mu=[mu], target_rho=[rho], target_lambda=[lam]))
        outs = self.run_sampler_batch(probs, num_iterations, quadrature_points, seed, device)
        shape = (len(sample.target_mu), len(sample.target_rho), len(sample.target_lambda))
        cat = lambda f: np.array([f(o) for o in outs])
        res = SamplerOutput(cat(lambda o: o.likelihood.ravel()[0]).reshape(shape), cat(lambda o: o.likelihoodSq.ravel()[0]).reshape(shape),
                            cat(lambda o: o.tmrca.ravel()[0]).reshape(shape),
                            np.stack([o.lineages_remaining.reshape(sample.T) for o in outs], axis=1).reshape((sample.T,) + shape),
                            cat(lambda o: o.ess.ravel()[0]).reshape(shape))
        res.n_histories = num_iterations
        res.n_events = sum(o.n_events for o in outs); res.n_pismc = sum(o.n_pismc for o in outs)
        res.seconds_device = sum(o.seconds_device for o in outs)
        return res


class DeviceSampler:
    """Staged API: tables + problem resident on one GPU; histories stay in HBM as torch tensors."""

    def __init__(self, problem: Problem, quadrature_points: int = 16, device: int = 0):
        self.L = load_library()
        self.problem = problem
        self.device = int(device)
        self._pk = _PackedProblem(problem)
        h = C.c_void_p()
        _check(self.L.coalgpu_create(C.byref(self._pk.struct), int(quadrature_points), self.device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.L.coalgpu_destroy(self._h)
            self._h = None

    __del__ = close

    def alloc(self, n: int, n_trace: int = 0, max_events: int = 0):
        import torch
        dev = torch.device("cuda", self.device)
        p = self.problem
        buf = dict(weights=torch.empty((n, p.G), dtype=torch.float64, device=dev),
                   tmrca=torch.empty((n,), dtype=torch.float64, device=dev),
                   lineages=torch.empty((n, p.T), dtype=torch.float64, device=dev),
                   nevents=torch.zeros((n,), dtype=torch.int32, device=dev),
                   partials=torch.empty((3 + p.T, p.G, 2), dtype=torch.float64, device=dev))
        if n_trace:
            buf["events"] = torch.zeros((n_trace, max_events, EVENT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
            buf["event_counts"] = torch.zeros((n_trace,), dtype=torch.int32, device=dev)
        return buf

    def sample(self, buf, seed: int, first: int, n: int, stream=None):
        """Draw histories [first, first+n) into buf (asynchronous on `stream` / torch's current stream)."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        ev = buf.get("events")
        _check(self.L.coalgpu_sample(self._h, int(seed), int(first), int(n), C.c_void_p(stream),
                                     buf["weights"].data_ptr(), buf["tmrca"].data_ptr(), buf["lineages"].data_ptr(),
                                     buf["nevents"].data_ptr(), ev.data_ptr() if ev is not None else None,
                                     buf["event_counts"].data_ptr() if ev is not None else None,
                                     ev.shape[0] if ev is not None else 0, ev.shape[1] if ev is not None else 0))

    def partials(self, buf, n: int, stream=None):
        """(max, sum exp) partials of the n histories in buf -> buf['partials'] ([3+T, G, 2], on device)."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        _check(self.L.coalgpu_partials(self._h, int(n), C.c_void_p(stream), buf["weights"].data_ptr(), buf["tmrca"].data_ptr(),
                                       buf["lineages"].data_ptr(), buf["partials"].data_ptr()))
        return buf["partials"]

    def finalize(self, partials_host: np.ndarray, n_histories: int = 0) -> SamplerOutput:
        p = self.problem
        hp = np.ascontiguousarray(partials_host, dtype=np.float64)
        arrs = _result_arrays(p)
        res = _CResult(_dp(arrs[0]), _dp(arrs[1]), _dp(arrs[2]), _dp(arrs[3]), _dp(arrs[4]), 0, 0, 0, 0.0)
        _check(self.L.coalgpu_finalize(p.G, p.T, hp.ctypes.data_as(C.c_void_p), C.byref(res)))
        out = _to_output(p, arrs)
        out.n_histories = n_histories
        return out

    def counters(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self.L.coalgpu_counters(self._h, C.byref(a), C.byref(b), C.byref(c)))
        d, e = C.c_uint64(), C.c_uint64()
        _check(self.L.coalgpu_work_counters(self._h, C.byref(d), C.byref(e)))
        out = dict(n_events=a.value, n_pismc=b.value, n_overflow=c.value, n_two_locus=d.value, n_classes=e.value)
        out.update(self.capacity_stats())
        return out

    def capacity_stats(self):
        """Two-pass capacity scheme: histories of the last sample() call deferred to the full type store, the two store
        sizes, the most distinct types any history needed so far, lanes per history of the first pass."""
        f = C.c_uint64(); k1, k2, pk, w = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(self.L.coalgpu_capacity_stats(self._h, C.byref(f), C.byref(k1), C.byref(k2), C.byref(pk), C.byref(w)))
        return dict(n_deferred=f.value, store_first_pass=k1.value, store_full=k2.value, peak_types=pk.value, lanes_per_history=w.value)

    def tune(self):
        """Re-size the reduced type store / launch geometry from the histories drawn so far (blocking; results never
        change, only occupancy). coalgpu_run_sampler does this itself after a pilot chunk."""
        _check(self.L.coalgpu_tune(self._h))
        return self.capacity_stats()

    def events(self, buf) -> List[np.ndarray]:
        """Traced event records per history (the integer history encoding)."""
        ev = buf["events"].cpu().numpy()
        cnt = buf["event_counts"].cpu().numpy()
        out = []
        for h in range(ev.shape[0]):
            rec = ev[h].reshape(-1).view(EVENT_DTYPE)
            out.append(rec[:min(int(cnt[h]), ev.shape[1])].copy())
        return out

    def tables(self, n_all: int, time_idx: int):
        p = self.problem
        Q = 16
        w = np.zeros(Q); y = np.zeros(Q); z = np.zeros((Q, Q)); EA = np.zeros((Q, p.LA + 1)); EB = np.zeros((Q, p.LB + 1))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        _check(self.L.coalgpu_tables(self._h, int(n_all), int(time_idx), vp(w), vp(y), vp(z), vp(EA), vp(EB)))
        return dict(w=w, y=y, z=z, EA=EA, EB=EB)

    def score_events(self, configs, chosen, epochs, max_cand=None):
        """The device event code on given configurations (coalgpu_score_events): configs = list of type lists
        [(gamma, word, count), ...] in slot order, chosen = list of (gamma, word), epochs = list of sampling epochs.
        Returns a list of dicts(values=[...], partner_words=[...]) in the kernel's candidate order, None where the
        configuration outgrew the type store."""
        p = self.problem
        if max_cand is None:
            max_cand = max(p.LA + p.LB + 4, max(p.LA, p.LB) + max((len(c) for c in configs), default=0) + 1)
        offs = np.zeros(len(configs) + 1, dtype=np.int32)
        g, w, c = [], [], []
        for q, cfg in enumerate(configs):
            for (gg, ww, cc) in cfg:
                g.append(gg); w.append(ww); c.append(cc)
            offs[q + 1] = len(g)
        g = np.asarray(g, dtype=np.uint8); w = np.asarray(w, dtype=np.uint64); c = np.asarray(c, dtype=np.uint32)
        xg = np.asarray([x[0] for x in chosen], dtype=np.uint8); xw = np.asarray([x[1] for x in chosen], dtype=np.uint64)
        ep = np.asarray(epochs, dtype=np.int32)
        pv = np.zeros((len(configs), max_cand)); pw = np.zeros((len(configs), max_cand), dtype=np.uint64)
        nc = np.zeros(len(configs), dtype=np.int32)
        vp = lambda x: x.ctypes.data_as(C.c_void_p)
        _check(self.L.coalgpu_score_events(self._h, len(configs), vp(offs), vp(g), vp(w), vp(c), vp(xg), vp(xw), vp(ep), int(max_cand),
                                           vp(pv), vp(pw), vp(nc)))
        return [None if nc[q] < 0 else dict(values=pv[q, :nc[q]].copy(), partner_words=pw[q, :nc[q]].copy()) for q in range(len(configs))]

    def pismc_classes(self, gamma, n_all, time_idx, class_offsets, counts, dA, dB) -> np.ndarray:
        """Batched piSMC from getDifferences output (ConditionalSamplingHMM.cpp:265-448) on the GPU."""
        g = np.ascontiguousarray(gamma, dtype=np.uint8); n = np.ascontiguousarray(n_all, dtype=np.int32)
        t = np.ascontiguousarray(time_idx, dtype=np.int32); o = np.ascontiguousarray(class_offsets, dtype=np.int64)
        c = np.ascontiguousarray(counts, dtype=np.int32); a = np.ascontiguousarray(dA, dtype=np.int32); b = np.ascontiguousarray(dB, dtype=np.int32)
        out = np.zeros(len(g), dtype=np.float64)
        vp = lambda x: x.ctypes.data_as(C.c_void_p)
        _check(self.L.coalgpu_pismc_classes(self._h, len(g), vp(g), vp(n), vp(t), vp(o), vp(c), vp(a), vp(b), vp(out)))
        return out


def fp64_peak_tflops(device: int = 0) -> float:
    """Measured FP64 FMA throughput of the device (TFLOP/s)."""
    v = C.c_double()
    _check(load_library().coalgpu_fp64_peak(int(device), C.byref(v)))
    return float(v.value)


def launch_count() -> int:
    return int(load_library().coalgpu_launch_count())


def transfer_bytes():
    """(host -> device, device -> host) bytes the drop-in entry points have copied in this process so far."""
    a, b = C.c_uint64(), C.c_uint64()
    load_library().coalgpu_transfer_bytes(C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def nccl_calls() -> int:
    """NCCL collectives issued by libcoalgpu.so itself (coalgpu_run_sampler_multi / _comm) in this process."""
    return int(load_library().coalgpu_nccl_calls())
