import ctypes as C
import glob
import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure). Built on demand from oracle/coal_oracle.cpp."""
    from oracle import oracle_py as O
    O.build()
    return O


@pytest.fixture(scope="session")
def cuda_lib():
    from coalescenator_b200 import build as B
    B.build_cuda()
    from coalescenator_b200 import sampler
    return sampler.load_library()


@pytest.fixture(scope="session")
def host_shim():
    """CPU harness around the product's host code / arithmetic cores (tests/host_shim.cpp)."""
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, "libcoalhost.so")
    srcs = [os.path.join(ROOT, "tests", "host_shim.cpp"), os.path.join(ROOT, "coalescenator_b200", "csrc", "coal_tables.cpp")]
    deps = srcs + [os.path.join(ROOT, "coalescenator_b200", "csrc", f) for f in ("coal_math.h", "coal_tables.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-o", lib] + srcs, check=True)
    L = C.CDLL(lib)
    L.hs_build.restype = C.c_void_p
    L.hs_build.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.hs_build_q.restype = C.c_void_p
    L.hs_build_q.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.hs_build_mode.restype = C.c_void_p
    L.hs_build_mode.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.hs_free.argtypes = [C.c_void_p]
    L.hs_tables.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5
    L.hs_pismc.restype = C.c_double
    L.hs_pismc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return L


def golden_names():
    return sorted(os.path.basename(f)[:-len(".coalprob")] for f in glob.glob(os.path.join(GOLDEN, "*.coalprob")))


def load_golden(name):
    from coalescenator_b200.problem import Problem
    p = Problem.load(os.path.join(GOLDEN, name + ".coalprob"))
    with gzip.open(os.path.join(GOLDEN, name + ".trace.gz"), "rt") as fh:
        trace = fh.read()
    with open(os.path.join(GOLDEN, name + ".result.json")) as fh:
        result = json.load(fh)
    return p, trace, result


def parse_pismc_records(trace, limit=None):
    """'P gamma n c LA LB nH (cnt dA dB)* pi' lines -> list of dicts."""
    out = []
    for line in trace.splitlines():
        if not line.startswith("P "):
            continue
        t = line.split()
        nH = int(t[6])
        cls = [int(x) for x in t[7:7 + 3 * nH]]
        out.append(dict(gamma="ABC".index(t[1]), n=int(t[2]), c=int(t[3]), cnt=cls[0::3], dA=cls[1::3], dB=cls[2::3], pi=float(t[7 + 3 * nH])))
        if limit and len(out) >= limit:
            break
    return out


def thetas_rrates(p):
    th = np.array([4 * p.driving_lambda * v * p.driving_mu for v in p.viral])
    rr = np.array([4 * p.driving_lambda * v * p.driving_rho for v in p.viral])
    return th, rr


def n_bound(p):
    return sum(c * (2 if g == 2 else 1) for tp in p.types for (g, _, c) in tp)
