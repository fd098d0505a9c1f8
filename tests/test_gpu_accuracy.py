"""The accuracy contract of the default (int8 digit-plane) path, GPU against the oracle.

The int8 path is fixed point relative to the row maximum of R.diag(effects) (include/gms_b200.h, GMS_MODE_INT8_TENSOR):
these tests feed it inputs that break plain 49-bit row-scaled digits - effect dynamic ranges of 1e6 and 1e12 with the
dominant SNP not segregating in some crosses, near-duplicate map positions (RoR close to singular), an all-zero trait,
one block longer than the 64-bit digit recombination can hold - and require the STRICT 1e-9 on every variance (no guard)
plus the guarded 1e-9 on covariances, in the default mode. The error bound of csrc/verify.cu must route the bad units to
the FP64 contraction; the statistics tell that it did."""
import numpy as np
import pytest

from oracle import predcrossvar_oracle as O
from tests.util import oracle_slab, oracle_slab_c, rel_err, strict_rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


def problem(seed, m_c, P, T, hom_parents=()):
    rng = np.random.default_rng(seed)
    m_c = np.asarray(m_c)
    m = int(m_c.sum())
    cM = np.concatenate([np.sort(rng.uniform(0, 110, k)) for k in m_c])
    chrom = np.repeat(np.arange(1, len(m_c) + 1), m_c)
    H = (rng.random((2 * P, m)) < rng.uniform(0.2, 0.8, m)).astype(np.uint8)
    add = rng.standard_normal((m, T))
    dom = 0.5 * rng.standard_normal((m, T))
    R = O.recombfreq_to_ld_decay(O.genmap2recombfreq(cM, chrom))
    return dict(m=m, m_c=m_c, cM=cM, chrom=chrom, H=H, add=add, dom=dom, R=R, rng=rng)


def run(g, s, model, add, dom, asub, sire, dam, mode="auto", tol=None):
    with g.CrossVarEngine(0) as eng:
        eng.set_genmap(s["cM"], s["m_c"])
        eng.set_haplotypes(s["H"])
        eng.set_mode(mode)
        if tol is not None:
            eng.set_int8_tolerance(tol)
        out, nseg = eng.predict(model, add, dom, asub, sire, dam)
        st = eng.stats()
    return out, nseg, st


@pytest.mark.parametrize("ratio", [1e6, 1e12])
@pytest.mark.parametrize("model", ["AD", "DirDom"])
def test_dominant_snp_not_segregating(ratio, model):
    """A few SNPs carry effects `ratio` times the rest. Parents 0-3 are homozygous there, parents 4-7 heterozygous: the
    crosses among 0-3 sum only small terms next to a row scale set by the huge ones."""
    import genomicmateselectr_b200 as g
    s = problem(41, [260, 190, 75], P=10, T=3)
    big = [17, 300, 301, 480]
    add, dom = s["add"].copy(), s["dom"].copy()
    add[big] *= ratio
    dom[big] *= ratio
    dom[big[1], 2] = 0.0
    H = s["H"]
    for p in range(4):
        H[2 * p + 1, big] = H[2 * p, big]             # homozygous at the dominant SNPs
    for p in range(4, 8):
        H[2 * p + 1, big] = 1 - H[2 * p, big]         # heterozygous
    asub = add + 0.3 * dom if model == "DirDom" else None
    sire = np.array([0, 0, 1, 2, 3, 0, 4, 5, 4, 8, 9, 2], np.int32)
    dam = np.array([1, 0, 2, 3, 3, 4, 5, 5, 7, 9, 0, 6], np.int32)
    want, wseg = oracle_slab(model, H, s["R"], add, dom, asub, sire, dam)
    out, nseg, st = run(g, s, model, add, dom, asub, sire, dam)
    assert st["last_mode"] == 2                                    # the int8 path ran ...
    assert np.array_equal(nseg, wseg)
    assert strict_rel_err(out, want) < TOL and rel_err(out, want) < TOL
    assert st["last_fallback_units"] > 0                           # ... and handed the units it cannot bound to FP64
    assert st["last_fallback_overflow"] == 0
    # the FP64 mode on the same input
    out64, _, st64 = run(g, s, model, add, dom, asub, sire, dam, mode="fp64")
    assert st64["last_mode"] == 0 and strict_rel_err(out64, want) < TOL and rel_err(out64, want) < TOL
    # with the bound switched off the plain digits are measurably worse on exactly these crosses (what the bound is for)
    if ratio >= 1e12:
        raw, _, st_raw = run(g, s, model, add, dom, asub, sire, dam, tol=0.0)
        assert st_raw["last_fallback_units"] == 0
        assert strict_rel_err(raw[:5], want[:5]) > 1e-7


def test_benign_effects_need_no_fallback():
    """Normal effects with a heavy tail: nothing fails the bound, so the int8 results are the int8 results."""
    import genomicmateselectr_b200 as g
    s = problem(43, [300, 200, 131], P=16, T=5)
    dom = s["dom"] * np.exp(s["rng"].normal(0, 1.0, (s["m"], 1)))
    rng = np.random.default_rng(1)
    sire = rng.integers(0, 16, 200).astype(np.int32)
    dam = rng.integers(0, 16, 200).astype(np.int32)
    want, wseg = oracle_slab("AD", s["H"], s["R"], s["add"], dom, None, sire, dam)
    out, nseg, st = run(g, s, "AD", s["add"], dom, None, sire, dam)
    assert st["last_mode"] == 2 and st["last_fallback_units"] == 0
    assert np.array_equal(nseg, wseg) and strict_rel_err(out, want) < TOL and rel_err(out, want) < TOL


def test_fallback_list_overflow_redoes_the_call_in_fp64():
    """More failing units than the in-stream lists hold (forced with an absurd tolerance): the call is redone in FP64."""
    import genomicmateselectr_b200 as g
    s = problem(44, [200, 150], P=12, T=2)
    rng = np.random.default_rng(2)
    N = 20000                                                       # > 16384 list entries
    sire = rng.integers(0, 12, N).astype(np.int32)
    dam = rng.integers(0, 12, N).astype(np.int32)
    out, nseg, st = run(g, s, "AD", s["add"], s["dom"], None, sire, dam, tol=1e-30)
    assert st["last_fallback_overflow"] == 1 and st["last_mode"] == 0
    pick = rng.integers(0, N, 40)
    want, wseg = oracle_slab("AD", s["H"], s["R"], s["add"], s["dom"], None, sire[pick], dam[pick])
    assert np.array_equal(nseg[pick], wseg) and rel_err(out[pick], want) < TOL
    ref, _, _ = run(g, s, "AD", s["add"], s["dom"], None, sire, dam, mode="fp64")
    assert np.array_equal(out, ref)                                 # bit-identical to the FP64 mode


def test_near_duplicate_map_positions():
    """Clusters of SNPs at (almost) the same position: R and RoR are numerically singular, covariances cancel."""
    import genomicmateselectr_b200 as g
    s = problem(45, [240, 160], P=8, T=3)
    cM = s["cM"].copy()
    for a in range(0, 240, 12):
        cM[a:a + 6] = cM[a] + 1e-9 * np.arange(6)
    cM[240:400] = np.sort(np.round(cM[240:400], 0))                # exact duplicates as well
    s["cM"] = cM
    s["R"] = O.recombfreq_to_ld_decay(O.genmap2recombfreq(cM, s["chrom"]))
    sire = np.array([0, 1, 2, 3, 4, 5, 6, 7, 3], np.int32)
    dam = np.array([1, 1, 5, 7, 0, 2, 6, 0, 4], np.int32)
    want, wseg = oracle_slab("AD", s["H"], s["R"], s["add"], s["dom"], None, sire, dam)
    for mode in ("auto", "fp64"):
        out, nseg, _ = run(g, s, "AD", s["add"], s["dom"], None, sire, dam, mode=mode)
        assert np.array_equal(nseg, wseg) and strict_rel_err(out, want) < TOL and rel_err(out, want) < TOL


@pytest.mark.parametrize("model", ["A", "AD", "DirDom"])
def test_all_zero_trait(model):
    import genomicmateselectr_b200 as g
    s = problem(46, [150, 131], P=8, T=4)
    add, dom = s["add"].copy(), s["dom"].copy()
    add[:, 1] = 0.0
    dom[:, 1] = 0.0
    dom[:, 3] = 0.0                                                 # a trait without dominance
    asub = add + 0.5 * dom if model == "DirDom" else None
    sire = np.array([0, 1, 2, 3, 7], np.int32)
    dam = np.array([4, 1, 6, 5, 0], np.int32)
    d = dom if model != "A" else None
    want, wseg = oracle_slab(model, s["H"], s["R"], add, d, asub, sire, dam)
    for mode in ("auto", "fp64", "map_scan"):
        out, nseg, _ = run(g, s, model, add, d, asub, sire, dam, mode=mode)
        assert np.array_equal(nseg, wseg) and rel_err(out, want) < TOL
        assert np.all(out[:, 1, :] == 0.0)                          # the zero trait's variance is exactly zero
        pairs = [(t, t) for t in range(4)] + [(a, b) for a in range(4) for b in range(a + 1, 4)]
        for i, (a, b) in enumerate(pairs):
            if 1 in (a, b):
                assert np.all(out[:, i, :] == 0.0)


def test_block_longer_than_the_digit_recombination_holds():
    """one block of more than 16384 SNPs, an all-ones matrix and equal effects: every digit sum has the same
    sign, and 16384 non-zero mask entries are where the exact 64-bit recombination of the digit sums could overflow.
    Such units must go to the FP64 path. All-ones R has a closed form: VarA = ((a.delta_S)^2 + (a.delta_D)^2) / 4,
    VarD = ((d.het_S)^2 + (d.het_D)^2 + 2 (d.xi)^2) / 16; the C oracle confirms it on the same input."""
    import genomicmateselectr_b200 as g
    m, P, T = 16640, 4, 2
    H = np.zeros((2 * P, m), np.uint8)
    H[0::2] = 1                                  # every parent heterozygous everywhere, same phase: delta = +1, xi = +1
    H[6, :700] = 0                               # parent 3: homozygous on a stretch
    add = np.ones((m, T))
    dom = np.ones((m, T))
    dom[:, 1] = 0.5
    R = np.ones((m, m))
    sire = np.array([0, 1, 3], np.int32)
    dam = np.array([1, 1, 2], np.int32)
    delta = H[0::2].astype(float) - H[1::2]
    het = np.abs(delta)
    want = np.zeros((3, 3, 2))
    pairs = [(0, 0), (1, 1), (0, 1)]
    for x, (s_, d_) in enumerate(zip(sire, dam)):
        xi = delta[s_] * delta[d_]
        for i, (a, b) in enumerate(pairs):
            want[x, i, 0] = ((add[:, a] @ delta[s_]) * (add[:, b] @ delta[s_]) + (add[:, a] @ delta[d_]) * (add[:, b] @ delta[d_])) / 4
            want[x, i, 1] = ((dom[:, a] @ het[s_]) * (dom[:, b] @ het[s_]) + (dom[:, a] @ het[d_]) * (dom[:, b] @ het[d_])
                             + 2 * (dom[:, a] @ xi) * (dom[:, b] @ xi)) / 16
    ref, rseg = oracle_slab_c("AD", H, add, dom, None, sire[:1], dam[:1], R=R)
    assert np.allclose(ref, want[:1], rtol=1e-12)
    with g.CrossVarEngine(0) as eng:
        assert eng.set_recomb_matrix(R, np.ones(m, int)) is True
        eng.set_haplotypes(H)
        out, nseg = eng.predict("AD", add, dom, None, sire, dam)
        st = eng.stats()
    assert st["last_mode"] == 2 and st["last_fallback_units"] > 0
    assert list(nseg) == [m, m, m] and rseg[0] == m
    assert np.allclose(out, want, rtol=1e-9, atol=0)


def test_error_bound_is_an_upper_bound():
    """The a-posteriori bound of verify.cu against the error actually made by the plain digits (verification off):
    the bound must never be smaller. Uses the dominant-SNP input, where the errors are large enough to measure."""
    import genomicmateselectr_b200 as g
    s = problem(47, [260, 190], P=8, T=2)
    dom = s["dom"].copy()
    dom[[40, 300]] *= 1e10
    H = s["H"]
    for p in range(4):
        H[2 * p + 1, [40, 300]] = H[2 * p, [40, 300]]
    sire = np.array([0, 1, 2, 0, 3], np.int32)
    dam = np.array([1, 2, 3, 3, 3], np.int32)
    want, _ = oracle_slab("AD", H, s["R"], s["add"], dom, None, sire, dam)
    raw, _, _ = run(g, s, "AD", s["add"], dom, None, sire, dam, tol=0.0)
    good, _, st = run(g, s, "AD", s["add"], dom, None, sire, dam)
    assert st["last_fallback_units"] > 0
    assert strict_rel_err(good, want) < TOL
    # the unverified result is off by far more than the default tolerance would have allowed: the bound had to fire
    assert strict_rel_err(raw, want) > 2.0 ** -31
