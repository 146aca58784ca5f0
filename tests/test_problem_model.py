"""The locus-pair data model (coalescenator_b200/problem.py): text round trip of the .coalprob format that the C++ host
driver, the reference drivers under oracle/ and the Python mirror share; bit layout of the packed haplotypes
(locus A site s = bit s, locus B site s = bit LA+s, the order of ProposalEngine.cpp:596-640); determinism of the
synthetic generator behind bench.py and the golden fixtures."""
import numpy as np

from conftest import golden_names, load_golden
from coalescenator_b200.problem import Problem, make_config, sub_problem, grid_cfg3, locus_pairs, tile_loci, synth_alignment


def test_golden_problems_round_trip():
    for name in golden_names():
        p, _, _ = load_golden(name)
        q = Problem.loads(p.dumps())
        assert q.dumps() == p.dumps()
        assert (q.LA, q.LB, q.T, q.G) == (p.LA, p.LB, p.T, p.G)
        assert q.types == p.types and q.times == p.times and q.viral == p.viral
        assert p.n_lineages() == sum(c for tp in p.types for _, _, c in tp)


def test_bit_layout():
    p = Problem(3, 2, [0.0], [100.0], 1.0, 1e-5, 1000.0, 5e-5, [1e-5], [5e-5], [1000.0], [[]])
    assert p.word(2, "10011") == 0b11001            # site 0 of locus A is bit 0; locus B starts at bit LA = 3
    assert p.word(0, "101") == 0b101 and p.word(1, "01") == 0b10 << 3
    for g, bits in ((2, "10011"), (0, "110"), (1, "10")):
        assert p.bits(g, p.word(g, bits)) == bits


def test_generator_is_deterministic_and_shaped_like_baseline_configs():
    a1, l1, p1 = make_config(2)
    a2, l2, p2 = make_config(2)
    assert l1 == l2 and p1 == p2 and all(np.array_equal(x, y) for x, y in zip(a1.seqs, a2.seqs))
    assert len(a1.seqs) == 5 and all(s.shape[0] == 10 for s in a1.seqs)           # cfg2: 5 time points x 10 sequences
    a4, l4, p4 = make_config(4)
    assert len(a4.seqs) == 10 and all(s.shape[0] == 20 for s in a4.seqs)          # cfg4: 10 time points x 20 sequences
    q = sub_problem(a1, l1, *p1[40])
    assert q.T == 5 and q.n_lineages() == 50 and q.LA >= 1 and q.LB >= 1 and q.LA + q.LB <= 64
    dist = p1[40][2]
    assert q.driving_rho == 5e-5 * (dist + 1) and q.target_rho == [5e-5 * (dist + 1)]     # main.cpp:260-271
    mu, rho, lam = grid_cfg3()
    assert (len(mu), len(rho), len(lam)) == (32, 1, 32)
    g = sub_problem(a1, l1, *p1[100], target_mu=mu, target_rho=rho, target_lambda=lam)
    assert g.G == 1024


def test_locus_tiling_and_pairs():
    aln = synth_alignment(3, 2, 4, 500, 40)
    loci = tile_loci(aln, 100)
    assert all(hi > lo for _, _, lo, hi in loci)
    pairs = locus_pairs(loci, min_dist=1, max_dist=150)
    assert pairs and all(i < j and 1 <= d <= 150 for i, j, d in pairs)
