"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (oracle/liboracle.so) and a runner for the
compiled reference (oracle/_ref/coalref). Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product path never does."""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODE_STRICT, MODE_CANONICAL = 0, 1


class _Problem(C.Structure):
    _fields_ = [("T", C.c_int32), ("LA", C.c_int32), ("LB", C.c_int32),
                ("times", C.POINTER(C.c_double)), ("viral", C.POINTER(C.c_double)),
                ("tau", C.c_double), ("driving_mu", C.c_double), ("driving_lambda", C.c_double), ("driving_rho", C.c_double),
                ("n_mu", C.c_int32), ("n_rho", C.c_int32), ("n_lambda", C.c_int32),
                ("mu", C.POINTER(C.c_double)), ("rho", C.POINTER(C.c_double)), ("lam", C.POINTER(C.c_double)),
                ("type_offsets", C.POINTER(C.c_int32)), ("type_gamma", C.POINTER(C.c_uint8)),
                ("type_words", C.POINTER(C.c_uint64)), ("type_counts", C.POINTER(C.c_uint32))]


EVENT_DTYPE = np.dtype([("chosen_word", "<u8"), ("aux_word", "<u8"), ("chosen_gamma", "u1"), ("kind", "u1"),
                        ("epoch", "u1"), ("site", "u1"), ("n_before", "<u4")])

_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.coalo_pismc_classes.restype = C.c_double
        L.coalo_uniform.restype = C.c_double
        L.coalo_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleProblem:
    """Keeps the numpy arrays alive behind a coalo_problem struct."""

    def __init__(self, p):
        self.p = p
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
        self.s = _Problem(p.T, p.LA, p.LB, _dp(self.times), _dp(self.viral), p.tau, p.driving_mu, p.driving_lambda,
                          p.driving_rho, len(self.mu), len(self.rho), len(self.lam), _dp(self.mu), _dp(self.rho), _dp(self.lam),
                          self.offs.ctypes.data_as(C.POINTER(C.c_int32)), self.g.ctypes.data_as(C.POINTER(C.c_uint8)),
                          self.w.ctypes.data_as(C.POINTER(C.c_uint64)), self.c.ctypes.data_as(C.POINTER(C.c_uint32)))

    @property
    def ref(self):
        return C.byref(self.s)


def tables(p, n_all, time_idx, Q=16):
    op = OracleProblem(p)
    w = np.zeros(Q); y = np.zeros(Q); z = np.zeros((Q, Q))
    EA = np.zeros((Q, p.LA + 1)); EB = np.zeros((Q, p.LB + 1))
    rc = lib().coalo_tables(op.ref, Q, int(n_all), int(time_idx), _dp(w), _dp(y), _dp(z), _dp(EA), _dp(EB))
    assert rc == 0, rc
    return dict(w=w, y=y, z=z, EA=EA, EB=EB)


def pismc_classes(op: OracleProblem, gamma, n_all, time_idx, counts, dA, dB, Q=16):
    cn = np.asarray(counts, dtype=np.int32); a = np.asarray(dA, dtype=np.int32); b = np.asarray(dB, dtype=np.int32)
    ip = lambda x: x.ctypes.data_as(C.POINTER(C.c_int32))
    return lib().coalo_pismc_classes(op.ref, Q, int(gamma), int(n_all), int(time_idx), len(cn), ip(cn), ip(a), ip(b))


def run(p, n, seed, mode=MODE_CANONICAL, first=0, Q=16, trace_path=None, trace_max=0, n_event_hist=0, max_events=4096):
    """Draw n histories. Returns dict(weights[n,G], tmrca[n], lineages[n,T], n_events[n], n_pismc[n], events)."""
    op = OracleProblem(p)
    G, T = p.G, p.T
    wts = np.zeros((n, G)); tm = np.zeros(n); lin = np.zeros((n, T))
    ne = np.zeros(n, dtype=np.int64); npi = np.zeros(n, dtype=np.int64)
    ev = np.zeros((max(n_event_hist, 1), max_events), dtype=EVENT_DTYPE)
    evc = np.zeros(max(n_event_hist, 1), dtype=np.int64)
    i64 = lambda x: x.ctypes.data_as(C.POINTER(C.c_int64))
    rc = lib().coalo_run(op.ref, Q, int(mode), C.c_uint64(seed), C.c_uint64(first), C.c_uint64(n), _dp(wts), _dp(tm), _dp(lin),
                         i64(ne), i64(npi), trace_path.encode() if trace_path else None, C.c_int64(trace_max),
                         ev.ctypes.data_as(C.c_void_p) if n_event_hist else None, C.c_int64(n_event_hist), C.c_int64(max_events), i64(evc))
    assert rc == 0, rc
    return dict(weights=wts, tmrca=tm, lineages=lin, n_events=ne, n_pismc=npi,
                events=[ev[h, :min(evc[h], max_events)].copy() for h in range(n_event_hist)])


def reduce(weights, tmrca, lineages):
    """SamplerOutput::sumSamples semantics. Returns dict(likelihood[G], likelihoodSq[G], tmrca[G], lineages[T,G])."""
    wts = np.ascontiguousarray(weights, dtype=np.float64); tm = np.ascontiguousarray(tmrca, dtype=np.float64)
    lin = np.ascontiguousarray(lineages, dtype=np.float64)
    n, G = wts.shape; T = lin.shape[1]
    a = np.zeros(G); b = np.zeros(G); c = np.zeros(G); d = np.zeros((T, G))
    rc = lib().coalo_reduce(C.c_int64(n), G, T, _dp(wts), _dp(tm), _dp(lin), _dp(a), _dp(b), _dp(c), _dp(d))
    assert rc == 0, rc
    return dict(likelihood=a, likelihoodSq=b, tmrca=c, lineages=d)


def set_score_dump(path, n_histories=0):
    """The next run() calls write one 'V' line per event of their first n_histories histories to `path` (None: off)."""
    lib().coalo_set_score_dump(path.encode() if path else None, C.c_int64(n_histories))


def parse_score_dump(path):
    """'V epoch gamma word ntypes (gamma word count)* ncand (label value)*' -> list of dicts."""
    out = []
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if not t or t[0] != "V":
                continue
            nt = int(t[4])
            cfg = [(int(t[5 + 3 * k]), int(t[6 + 3 * k], 16), int(t[7 + 3 * k])) for k in range(nt)]
            o = 5 + 3 * nt
            nc = int(t[o])
            cand = [(t[o + 1 + 2 * k], float(t[o + 2 + 2 * k])) for k in range(nc)]
            out.append(dict(epoch=int(t[1]), chosen=(int(t[2]), int(t[3], 16)), config=cfg, cand=cand))
    return out


def score_event(p, config, chosen, epoch, Q=16, max_cand=512):
    """Candidate vector of one event in CANONICAL mode (coalo_score_event): list of (label, value)."""
    op = OracleProblem(p)
    g = np.asarray([c[0] for c in config], dtype=np.uint8); w = np.asarray([c[1] for c in config], dtype=np.uint64)
    c = np.asarray([c[2] for c in config], dtype=np.uint32)
    vals = np.zeros(max_cand); labels = C.create_string_buffer(24 * max_cand)
    vp = lambda x: x.ctypes.data_as(C.c_void_p)
    n = lib().coalo_score_event(op.ref, Q, int(epoch), len(config), vp(g), vp(w), vp(c), int(chosen[0]), C.c_uint64(int(chosen[1])),
                                max_cand, vp(vals), labels)
    assert n >= 0, n
    return [(labels.raw[24 * k:24 * k + 24].split(b"\0")[0].decode(), float(vals[k])) for k in range(n)]


def uniform(seed, hist, k):
    return lib().coalo_uniform(seed, hist, k)


# ---- compiled reference (oracle/_ref) ---------------------------------------------------------------

def ref_binary(trace=False):
    path = os.path.join(HERE, "_ref", "coalref_trace" if trace else "coalref")
    return path if os.path.exists(path) else None


def run_reference(p, n, seed, threads=1, trace_path=None, trace_max=0, timeout=None):
    """Run the compiled, unmodified reference (ImportanceSampler::runSampler) on problem p."""
    exe = ref_binary(trace=trace_path is not None)
    if exe is None:
        raise FileNotFoundError("oracle/_ref not built (needs /root/reference; run `make -C oracle ref`)")
    with tempfile.NamedTemporaryFile("w", suffix=".coalprob", delete=False) as fh:
        fh.write(p.dumps())
        prob = fh.name
    try:
        cmd = [exe, prob, str(n), str(threads), str(seed)]
        if trace_path:
            cmd += [trace_path, str(trace_max)]
        out = subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=timeout).stdout
    finally:
        os.unlink(prob)
    return json.loads(out)
