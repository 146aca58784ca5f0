"""Pins the CPU oracle (oracle/coal_oracle.cpp, STRICT mode) to the reference.

The golden fixtures were produced by running the compiled, UNMODIFIED reference
(tests/golden/make_golden.py -> oracle/_ref/coalref{,_trace}); the reference ships no tests or
golden vectors of its own. The oracle must reproduce them bit for bit:
every piSMC input/output, every lineage choice, every state change, every per-history weight,
and the final SamplerOutput grids."""
import os

import numpy as np
import pytest

from conftest import golden_names, load_golden


@pytest.mark.parametrize("name", golden_names())
def test_trace_bit_identical(oracle, name, tmp_path):
    p, trace, result = load_golden(name)
    n_trace = trace.count("\nB ") + (1 if trace.startswith("B ") else 0)
    out = tmp_path / "oracle.trace"
    oracle.run(p, n_trace, result["seed"], mode=oracle.MODE_STRICT, trace_path=str(out), trace_max=n_trace)
    mine = out.read_text()
    if mine != trace:
        a, b = mine.splitlines(), trace.splitlines()
        for i, (x, y) in enumerate(zip(a, b)):
            assert x == y, f"{name}: first differing trace line {i}:\n oracle   : {x}\n reference: {y}"
        assert len(a) == len(b)


@pytest.mark.parametrize("name", golden_names())
def test_run_sampler_output_bit_identical(oracle, name):
    p, _, result = load_golden(name)
    r = oracle.run(p, result["n_samples"], result["seed"], mode=oracle.MODE_STRICT)
    red = oracle.reduce(r["weights"], r["tmrca"], r["lineages"])
    assert np.array_equal(red["likelihood"], np.array(result["likelihood"]))
    assert np.array_equal(red["likelihoodSq"], np.array(result["likelihoodSq"]))
    assert np.array_equal(red["tmrca"], np.array(result["tmrca"]))
    for l in range(p.T):
        assert np.array_equal(red["lineages"][l], np.array(result["lineages_remaining"][str(l)]))


def test_live_reference_if_built(oracle, tmp_path):
    """When oracle/_ref exists (it is built wherever /root/reference is mounted) compare a fresh seed too."""
    if oracle.ref_binary(trace=True) is None:
        pytest.skip("oracle/_ref not built here")
    p, _, _ = load_golden("missing_pair3")
    ref_t, ora_t = tmp_path / "ref.trace", tmp_path / "ora.trace"
    ref = oracle.run_reference(p, 5, 12345, trace_path=str(ref_t), trace_max=5)
    r = oracle.run(p, 5, 12345, mode=oracle.MODE_STRICT, trace_path=str(ora_t), trace_max=5)
    assert ref_t.read_text() == ora_t.read_text()
    red = oracle.reduce(r["weights"], r["tmrca"], r["lineages"])
    assert np.array_equal(red["likelihood"], np.array(ref["likelihood"]))


def test_canonical_mode_agrees_statistically(oracle):
    """CANONICAL mode (Philox + slot order + order-free class representative) is the same estimator:
    its log-likelihood agrees with STRICT mode within Monte-Carlo error."""
    p, _, _ = load_golden("cfg1_pair0")
    est = []
    for mode in (oracle.MODE_STRICT, oracle.MODE_CANONICAL):
        r = oracle.run(p, 3000, 2024, mode=mode)
        lw = r["weights"][:, 0]
        m = lw.max()
        w = np.exp(lw - m)
        est.append((np.log(w.mean()) + m, np.sqrt((w ** 2).sum()) / w.sum()))
    (a, sa), (b, sb) = est
    assert abs(a - b) < 5 * np.hypot(sa, sb) + 0.05, (est)


@pytest.mark.parametrize("name,n", [("cfg1_pair0", 4000), ("missing_pair3", 6000)])
def test_canonical_mode_against_20000_reference_histories(oracle, name, n):
    """The same link against a large run of the compiled reference itself (tests/golden/<name>.big_result.json, 20,000
    histories, effective sample size 237 / 569): CANONICAL mode's log mean weight within 4 combined standard errors."""
    import json, os
    from conftest import GOLDEN
    ref = json.load(open(os.path.join(GOLDEN, name + ".big_result.json")))
    p, _, _ = load_golden(name)
    r = oracle.run(p, n, 31337, mode=oracle.MODE_CANONICAL)
    lw = r["weights"][:, 0]
    m = lw.max()
    w = np.exp(lw - m)
    mine, se_mine = np.log(w.mean()) + m, np.sqrt(max((w ** 2).sum() / w.sum() ** 2 - 1.0 / n, 0))
    rl, rl2, n_ref = ref["likelihood"][0], ref["likelihoodSq"][0], ref["n_samples"]
    ref_like, ref_se = rl - np.log(n_ref), np.sqrt(max(np.exp(rl2 - 2 * rl) - 1.0 / n_ref, 0))
    z = (mine - ref_like) / np.hypot(se_mine, ref_se)
    print(name, "log mean w: CANONICAL", mine, "reference", ref_like, "z", z)
    assert abs(z) < 4.0


def test_philox_known_answer(oracle):
    # Philox4x32-10 known-answer test from the Random123 distribution (kat_vectors):
    # counter = key = 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    u0 = oracle.uniform(0, 0, 0)
    lo, hi = 0x6627e8d5, 0xe169c58d
    # 52 bits per draw: (bits + 0.5) / 2^52 is exact in double, never 0 and never 1 (53 bits + 0.5 rounds to 1.0)
    assert u0 == (((hi << 32 | lo) >> 12) + 0.5) / 2 ** 52
    u1 = oracle.uniform(0, 0, 1)
    lo, hi = 0xbc57ac4c, 0x9b00dbd8
    assert u1 == (((hi << 32 | lo) >> 12) + 0.5) / 2 ** 52
    assert (2 ** 52 - 1 + 0.5) / 2 ** 52 < 1.0 and 0.5 / 2 ** 52 > 0.0
