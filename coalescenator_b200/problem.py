"""Host-side data model for one locus-pair problem and the synthetic data generator.

`Problem` mirrors what the reference hands to `ImportanceSampler::runSampler`
(/root/reference/src/ImportanceSampler.cpp:110): a `Sample` (Sample.hpp:44-63: per-time-point
A/B/C type maps, sampling times, viral loads, locus lengths) plus `ModelParameters`
(Utils.hpp:157-166: driving values, target grids, generation time tau).

Haplotypes are bit-packed: site s of locus A is bit s, site s of locus B is bit LA+s of one 64-bit
word (the reference stores `vector<bool>` with locus A first, ProposalEngine.cpp:596-640).
The text file format written here is shared by the CPU oracle, the compiled reference driver
(oracle/ref_driver.cpp) and the C++ host driver.
"""
from __future__ import annotations

import dataclasses
import itertools
from typing import List, Sequence, Tuple

import numpy as np

GAMMA_A, GAMMA_B, GAMMA_C = 0, 1, 2
GAMMA_NAMES = "ABC"


@dataclasses.dataclass
class Problem:
    LA: int
    LB: int
    times: List[float]
    viral: List[float]
    tau: float
    driving_mu: float
    driving_lambda: float
    driving_rho: float
    target_mu: List[float]
    target_rho: List[float]
    target_lambda: List[float]
    # per time point: list of (gamma in {0,1,2}, packed word, count), in insertion order
    types: List[List[Tuple[int, int, int]]]
    name: str = ""

    @property
    def T(self) -> int:
        return len(self.times)

    @property
    def G(self) -> int:
        return len(self.target_mu) * len(self.target_rho) * len(self.target_lambda)

    def n_lineages(self) -> int:
        return sum(c for tp in self.types for (_, _, c) in tp)

    # ---- bit helpers -------------------------------------------------------------------------
    def bits(self, gamma: int, word: int) -> str:
        if gamma == GAMMA_A:
            return "".join("1" if (word >> s) & 1 else "0" for s in range(self.LA))
        if gamma == GAMMA_B:
            return "".join("1" if (word >> (self.LA + s)) & 1 else "0" for s in range(self.LB))
        return "".join("1" if (word >> s) & 1 else "0" for s in range(self.LA + self.LB))

    def word(self, gamma: int, bits: str) -> int:
        off = self.LA if gamma == GAMMA_B else 0
        w = 0
        for s, ch in enumerate(bits):
            if ch == "1":
                w |= 1 << (off + s)
        return w

    # ---- text format -------------------------------------------------------------------------
    def dumps(self) -> str:
        f = lambda v: " ".join(repr(float(x)) for x in v)
        out = ["coalprob 1", f"T {self.T} LA {self.LA} LB {self.LB}",
               "times " + f(self.times), "viral " + f(self.viral), f"tau {float(self.tau)!r}",
               f"driving {float(self.driving_mu)!r} {float(self.driving_lambda)!r} {float(self.driving_rho)!r}",
               f"mu {len(self.target_mu)} " + f(self.target_mu),
               f"rho {len(self.target_rho)} " + f(self.target_rho),
               f"lambda {len(self.target_lambda)} " + f(self.target_lambda)]
        for t, tp in enumerate(self.types):
            out.append(f"types {t} {len(tp)}")
            for (g, w, c) in tp:
                out.append(f"{GAMMA_NAMES[g]} {self.bits(g, w)} {c}")
        return "\n".join(out) + "\n"

    def save(self, path) -> None:
        with open(path, "w") as fh:
            fh.write(self.dumps())

    @staticmethod
    def loads(text: str, name: str = "") -> "Problem":
        tok = text.split()
        pos = 0

        def take(n=1):
            nonlocal pos
            v = tok[pos:pos + n]
            pos += n
            return v

        assert take(2) == ["coalprob", "1"], "bad header"
        _, T, _, LA, _, LB = take(6)
        T, LA, LB = int(T), int(LA), int(LB)
        assert take()[0] == "times"
        times = [float(x) for x in take(T)]
        assert take()[0] == "viral"
        viral = [float(x) for x in take(T)]
        assert take()[0] == "tau"
        tau = float(take()[0])
        assert take()[0] == "driving"
        dmu, dlam, drho = (float(x) for x in take(3))
        grids = {}
        for key in ("mu", "rho", "lambda"):
            assert take()[0] == key
            n = int(take()[0])
            grids[key] = [float(x) for x in take(n)]
        p = Problem(LA, LB, times, viral, tau, dmu, dlam, drho, grids["mu"], grids["rho"], grids["lambda"], [], name)
        for t in range(T):
            k, tt, n = take(3)
            assert k == "types" and int(tt) == t
            tp = []
            for _ in range(int(n)):
                g, bits, c = take(3)
                gi = GAMMA_NAMES.index(g)
                tp.append((gi, p.word(gi, bits), int(c)))
            p.types.append(tp)
        return p

    @staticmethod
    def load(path) -> "Problem":
        with open(path) as fh:
            return Problem.loads(fh.read(), name=str(path))


# ---- synthetic serially-sampled alignments ---------------------------------------

@dataclasses.dataclass
class Alignment:
    """Bi-allelic segregating sites of a synthetic alignment: what is left of a FASTA set after the
    reference's site filters (main.cpp:119-121). seqs[t] is an (n_t, S) 0/1 array; missing[t] marks,
    per sequence, a (left, right) pair of site counts that are missing at the ends (the reference's
    left/right '-' runs, Sample.cpp:205-251)."""
    length: int
    positions: np.ndarray            # (S,) sorted 0-based nucleotide positions
    seqs: List[np.ndarray]
    missing: List[np.ndarray]        # (n_t, 2) ints: sites < left or >= S-right are missing
    times: List[float]
    viral: List[float]


def synth_alignment(seed: int, n_times: int, seqs_per_time: int, length: int, n_seg: int,
                    founders: int = 6, p_founder: float = 0.3, p_flip: float = 0.05,
                    dt: float = 50.0, viral: float = 100.0, p_missing: float = 0.0) -> Alignment:
    rng = np.random.default_rng(seed)
    pos = np.sort(rng.choice(length, size=n_seg, replace=False))
    fnd = (rng.random((founders, n_seg)) < p_founder).astype(np.uint8)
    seqs, missing = [], []
    for _ in range(n_times):
        pick = rng.integers(0, founders, size=seqs_per_time)
        s = fnd[pick] ^ (rng.random((seqs_per_time, n_seg)) < p_flip).astype(np.uint8)
        seqs.append(s)
        m = np.zeros((seqs_per_time, 2), dtype=np.int64)
        if p_missing > 0:
            has = rng.random((seqs_per_time, 2)) < p_missing
            m = has * rng.integers(1, max(2, n_seg // 2), size=(seqs_per_time, 2))
        missing.append(m)
    # make every site bi-allelic over the whole sample
    allseq = np.concatenate(seqs, axis=0)
    mono = np.where((allseq.min(axis=0) == allseq.max(axis=0)))[0]
    for k, site in enumerate(mono):
        t = k % n_times
        seqs[t][k % seqs_per_time, site] ^= 1
    return Alignment(length, pos, seqs, missing, [dt * t for t in range(n_times)], [viral] * n_times)


def tile_loci(aln: Alignment, loci_size: int) -> List[Tuple[int, int, int, int]]:
    """Tile [0, length) into windows of loci_size nucleotides (Sample.cpp:386-411) and keep those that
    contain >= 1 segregating site (Sample.cpp:537-557). Returns (first_nt, last_nt, site_lo, site_hi)
    with 1-based inclusive nucleotide coordinates and a half-open range into aln.positions."""
    loci = []
    left = 1
    while left <= aln.length:
        right = min(left + loci_size - 1, aln.length)
        if right - left + 1 < loci_size and loci and right == aln.length and left + loci_size - 1 > aln.length:
            pass
        lo = int(np.searchsorted(aln.positions, left - 1, side="left"))
        hi = int(np.searchsorted(aln.positions, right - 1, side="right"))
        if hi > lo:
            loci.append((left, right, lo, hi))
        left = right + 1
    return loci


def locus_pairs(loci, min_dist: int = 1, max_dist: int = 400):
    """Pairs (i, j, distance) of loci whose gap in nucleotides is within [min_dist, max_dist]
    (main.cpp:242-252)."""
    out = []
    for i, j in itertools.combinations(range(len(loci)), 2):
        d = loci[j][0] - loci[i][1] - 1
        if min_dist <= d <= max_dist:
            out.append((i, j, d))
    return out


def sub_problem(aln: Alignment, loci, i: int, j: int, distance: int, *, tau: float = 1.0,
                driving=(2.5e-5, 1000.0, 5e-5), target_mu=(2.5e-5,), target_rho=(5e-5,),
                target_lambda=(1000.0,), max_bits: int = 64) -> Problem:
    """Extract the A/B/C type maps of one locus pair (Sample.cpp:587-692) and scale rho by
    distance+1 (main.cpp:260-271)."""
    _, _, a_lo, a_hi = loci[i]
    _, _, b_lo, b_hi = loci[j]
    LA, LB = a_hi - a_lo, b_hi - b_lo
    assert LA + LB <= max_bits, "locus pair too long for one packed word"
    S = len(aln.positions)
    types, times, viral = [], [], []
    for t, s in enumerate(aln.seqs):
        order, counts = [], {}
        for r in range(s.shape[0]):
            ml, mr = aln.missing[t][r]
            okA = a_lo >= ml and a_hi <= S - mr
            okB = b_lo >= ml and b_hi <= S - mr
            if not okA and not okB:
                continue
            w = 0
            if okA:
                for k in range(LA):
                    w |= int(s[r, a_lo + k]) << k
            if okB:
                for k in range(LB):
                    w |= int(s[r, b_lo + k]) << (LA + k)
            g = GAMMA_C if (okA and okB) else (GAMMA_A if okA else GAMMA_B)
            if (g, w) not in counts:
                counts[(g, w)] = 0
                order.append((g, w))
            counts[(g, w)] += 1
        if order:
            types.append([(g, w, counts[(g, w)]) for (g, w) in order])
            times.append(aln.times[t])
            viral.append(aln.viral[t])
    scale = distance + 1
    return Problem(LA, LB, times, viral, tau, driving[0], driving[1], driving[2] * scale,
                   list(target_mu), [r * scale for r in target_rho], list(target_lambda), types,
                   name=f"pair{i}-{j}_d{distance}")


def make_config(cfg: int, seed: int = 1, **kw):
    """The BASELINE.json configurations as synthetic inputs. Returns (alignment, loci, pairs)."""
    if cfg == 1:
        aln = synth_alignment(seed, 2, 5, 1000, 30, **kw)
    elif cfg in (2, 3, 5):
        aln = synth_alignment(seed, 5, 10, 10000, 300, **kw)
    elif cfg == 4:
        aln = synth_alignment(seed, 10, 20, 100000, 3000, **kw)
    else:
        raise ValueError(cfg)
    loci = tile_loci(aln, 100)
    return aln, loci, locus_pairs(loci)


def grid_cfg3():
    """32 x 1 x 32 (mu, rho, lambda) grid of BASELINE.json config 3."""
    mu = list(np.exp(np.linspace(np.log(1e-6), np.log(1e-4), 32)))
    lam = list(np.exp(np.linspace(np.log(250.0), np.log(4000.0), 32)))
    return mu, [5e-5], lam
