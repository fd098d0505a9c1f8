"""Algebraic properties of the reference formulation (checked on the CPU oracle) that the CUDA kernels rely on
(SURVEY.md section 7, step 1): delta/het form, symmetry, self-cross = 2 x gametic, HapA/HapB swap,
dropping non-segregating SNPs, additivity over chromosomes, prefix-sum form of a map-derived matrix."""
import numpy as np
import pytest

from oracle import predcrossvar_oracle as O
from tests.util import oracle_slab


def problem(seed, m_c=(40, 25), P=6, T=3):
    rng = np.random.default_rng(seed)
    m_c = np.asarray(m_c)
    m = int(m_c.sum())
    cM = np.concatenate([np.sort(rng.uniform(0, 100, k)) for k in m_c])
    chrom = np.repeat(np.arange(1, len(m_c) + 1), m_c)
    H = (rng.random((2 * P, m)) < rng.uniform(0.1, 0.9, m)).astype(np.uint8)
    R = O.recombfreq_to_ld_decay(O.genmap2recombfreq(cM, chrom))
    return dict(m=m, m_c=m_c, cM=cM, chrom=chrom, H=H, R=R, add=rng.standard_normal((m, T)), dom=rng.standard_normal((m, T)))


@pytest.mark.parametrize("seed", range(5))
def test_delta_het_xi_form_of_the_quadratic_forms(seed):
    """VarA = 1/4 (Q_S + Q_D), VarD = 1/16 (P_S + P_D + 2 X) with masks from the haplotype bits."""
    p = problem(seed)
    sire, dam = [0, 1, 2], [3, 1, 5]
    want, _ = oracle_slab("AD", p["H"], p["R"], p["add"], p["dom"], None, sire, dam)
    R2 = p["R"] * p["R"]
    T = p["add"].shape[1]
    pairs = [(t, t) for t in range(T)] + [(a, b) for a in range(T) for b in range(a + 1, T)]
    for x, (s, d) in enumerate(zip(sire, dam)):
        ds = p["H"][2 * s].astype(float) - p["H"][2 * s + 1]
        dd = p["H"][2 * d].astype(float) - p["H"][2 * d + 1]
        xi = ds * dd
        for i, (a, b) in enumerate(pairs):
            q = lambda M, e, mk: (e[:, a] * mk) @ M @ (e[:, b] * mk)
            va = 0.25 * (q(p["R"], p["add"], ds) + q(p["R"], p["add"], dd))
            vd = 0.0625 * (q(R2, p["dom"], ds * ds) + q(R2, p["dom"], dd * dd) + 2 * q(R2, p["dom"], xi))
            assert va == pytest.approx(want[x, i, 0], rel=1e-11, abs=1e-12)
            assert vd == pytest.approx(want[x, i, 1], rel=1e-11, abs=1e-12)


def test_symmetry_selfcross_and_haplotype_swap():
    p = problem(11)
    a, _ = oracle_slab("AD", p["H"], p["R"], p["add"], p["dom"], None, [0, 2], [1, 2])
    b, _ = oracle_slab("AD", p["H"], p["R"], p["add"], p["dom"], None, [1, 2], [0, 2])
    assert np.allclose(a, b, rtol=1e-12)                                   # sire <-> dam
    H2 = p["H"].copy()
    H2[[0, 1]] = H2[[1, 0]]                                                # HapA <-> HapB of parent 0
    c, _ = oracle_slab("AD", H2, p["R"], p["add"], p["dom"], None, [0, 2], [1, 2])
    assert np.allclose(a, c, rtol=1e-12)
    X = p["H"][[4, 5]].astype(float)
    D = O.calcGameticLD(X, p["R"])
    assert np.array_equal(O.calcCrossLD(X, X, p["R"]), 2 * D)              # self cross = 2 x gametic LD


def test_nonsegregating_snps_do_not_matter_and_chromosomes_add_up():
    p = problem(12)
    H = p["H"].copy()
    H[:, 5] = 1
    H[:, 17] = 0                                                            # fixed everywhere
    full, nseg = oracle_slab("AD", H, p["R"], p["add"], p["dom"], None, [0, 1], [2, 3])
    keep = np.ones(p["m"], bool)
    keep[[5, 17]] = False
    sub, nseg2 = oracle_slab("AD", H[:, keep], p["R"][np.ix_(keep, keep)], p["add"][keep], p["dom"][keep], None, [0, 1], [2, 3])
    assert np.array_equal(nseg, nseg2) and np.allclose(full, sub, rtol=1e-12)
    tot = 0
    for c in (1, 2):
        k = p["chrom"] == c
        o, _ = oracle_slab("AD", H[:, k], p["R"][np.ix_(k, k)], p["add"][k], p["dom"][k], None, [0, 1], [2, 3])
        tot = tot + o
    assert np.allclose(full, tot, rtol=1e-12)


def test_prefix_sum_form_of_a_map_derived_matrix():
    """w' R v = sum_j [w_j v_j + a_j w_j G_j(v) + a_j v_j G_j(w)], G_j(v) = sum_{k<j} b_k v_k (csrc/scan.cu)."""
    rng = np.random.default_rng(3)
    x = np.sort(rng.uniform(0, 150, 60))
    R = O.recombfreq_to_ld_decay(O.genmap2recombfreq(x, np.ones(60)))
    for c, M in ((0.02, R), (0.04, R * R)):
        a, b = np.exp(-c * (x - x[0])), np.exp(c * (x - x[0]))
        w, v = rng.standard_normal(60), rng.standard_normal(60)
        Gv = np.concatenate([[0], np.cumsum(b * v)[:-1]])
        Gw = np.concatenate([[0], np.cumsum(b * w)[:-1]])
        assert np.sum(w * v + a * w * Gv + a * v * Gw) == pytest.approx(w @ M @ v, rel=1e-12)


def test_int8_digit_planes_are_an_error_free_split():
    """Row-scaled balanced radix-256 digits (csrc/i8digits.cuh): with rowmax <= 0.996 * 2^E and w = rn(G 2^(47 - E)),
    w = sum_i g_i 256^(5 - i), g_i in [-128, 127], |w 2^(E - 47) - G| <= 2^(E - 48); ternary dot products of the digit
    planes recombine to the exact integer dot product."""
    rng = np.random.default_rng(4)
    G = rng.standard_normal((20, 300)) * np.exp(rng.normal(0, 3, (20, 1)))
    f, E = np.frexp(np.abs(G).max(axis=1))
    E = E + (f > 0.996)
    w = np.rint(G / np.ldexp(1.0, E)[:, None] * 2.0 ** 47).astype(np.int64)
    assert np.all(np.abs(w) <= 127 * 0x010101010101)
    digits, r = [], w.copy()
    for _ in range(6):
        g = ((r + 128) & 255) - 128
        digits.append(g)
        r = (r - g) >> 8
    assert np.all(r == 0) and all(np.all((g >= -128) & (g <= 127)) for g in digits)
    digits = digits[::-1]                                     # most significant first
    rec = sum(g.astype(np.int64) << (8 * (5 - i)) for i, g in enumerate(digits))
    assert np.array_equal(rec, w)
    assert np.all(np.abs(rec * np.ldexp(1.0, E - 47)[:, None] - G) <= np.ldexp(1.0, E - 48)[:, None])
    xi = rng.integers(-1, 2, 300)
    S = [g @ xi for g in digits]                              # what the tensor cores accumulate, one plane at a time
    W = [sum(int(S[i][j]) << (8 * (5 - i)) for i in range(6)) for j in range(20)]
    assert W == [int(v) for v in (w @ xi)]
