"""Pin the CPU oracle to every known answer the reference publishes for the hot path:
the printed vignette outputs (docs/articles/*.html), 3-4 significant figures."""
import numpy as np
import pytest

from oracle import predcrossvar_oracle as O


def sig(x, n):
    """Round to n significant figures the way R's tibble printing does."""
    if x == 0:
        return 0.0
    from math import floor, log10
    return round(x, n - 1 - floor(log10(abs(x))))


def run(b, model, add, dom=None):
    crosses = O.crosses2predict(b.PARENTS)
    return O.predCrossVars(crosses, model, add, dom, haploMat=b.haploMat,
                           hap_rownames=b.hap_rownames, recombFreqMat=b.recombFreqMat)


def lookup(res, sire, dam, t1, t2, pred_of):
    for r in res:
        if r["sireID"] == sire and r["damID"] == dam:
            for a, c, p, v in r["predVars"]:
                if (a, c, p) == (t1, t2, pred_of):
                    return v
    raise KeyError((sire, dam, t1, t2, pred_of))


def test_crosses2predict_order():
    # docs/articles/start_here.html:281-288
    c = O.crosses2predict(["49", "153", "74", "228", "146"])
    assert len(c) == 15
    assert c[:6] == [("49", "49"), ("49", "153"), ("49", "74"), ("49", "228"), ("49", "146"), ("153", "153")]


def test_vignette_A(bundled):
    b = bundled
    res = run(b, "A", b.effdict(b.effA))
    assert len(res) == 15
    # Nsegsnps, docs/articles/start_here.html:349 (tidy rows are 3 per cross)
    assert [r["Nsegsnps"] for r in res[:4]] == [22, 46, 48, 53]
    nseg = {(r["sireID"], r["damID"]): r["Nsegsnps"] for r in res}
    assert nseg[("153", "153")] == 33 and nseg[("153", "228")] == 55      # :404,406
    # raw predVars order and values, :382-385: (T1,T1), (T2,T2), (T1,T2) then next cross
    flat = [(a, c, v) for r in res for a, c, _, v in r["predVars"]]
    assert [(a, c) for a, c, _ in flat[:3]] == [("Trait1", "Trait1"), ("Trait2", "Trait2"), ("Trait1", "Trait2")]
    printed = [707, 1519, -350, 1079, 1813]
    assert [sig(v, len(str(abs(p)))) for (_, _, v), p in zip(flat, printed)] == printed
    # tidy predVar :333-342
    assert sig(lookup(res, "49", "74", "Trait1", "Trait1", "VarA"), 3) == 974
    assert sig(lookup(res, "49", "74", "Trait2", "Trait2", "VarA"), 4) == 1510
    # SELIND variance with SIwts (.75,.25): 361, 454, 458, 536 (:333-342); 546, 628 (:404-406)
    SI = {"Trait1": 0.75, "Trait2": 0.25}
    want = {("49", "49"): 361, ("49", "153"): 454, ("49", "74"): 458, ("49", "228"): 536,
            ("153", "153"): 546, ("153", "228"): 628}
    for r in res:
        key = (r["sireID"], r["damID"])
        if key in want:
            rows = [(a, c, v) for a, c, _, v in r["predVars"]]
            assert sig(O.selind_var(rows, b.traits, SI), 3) == want[key]


def test_vignette_AD(bundled):
    b = bundled
    res = run(b, "AD", b.effdict(b.effAD_add), b.effdict(b.effAD_dom))
    g = lambda s, d, a, c, p: lookup(res, s, d, a, c, p)
    # docs/articles/non_additive_models.html:624-631 (VarA is relabelled VarBV, VarD -> VarDD)
    assert sig(g("146", "146", "Trait1", "Trait1", "VarA"), 4) == 1232
    assert sig(g("146", "146", "Trait1", "Trait2", "VarA"), 3) == -820
    assert sig(g("146", "146", "Trait2", "Trait2", "VarA"), 4) == 1976
    assert sig(g("146", "146", "Trait1", "Trait1", "VarD"), 3) == 9.97
    assert sig(g("146", "146", "Trait1", "Trait2", "VarD"), 3) == 9.04
    assert sig(g("146", "146", "Trait2", "Trait2", "VarD"), 3) == 51.3
    assert sig(g("153", "146", "Trait1", "Trait1", "VarA"), 4) == 1360
    assert sig(g("153", "146", "Trait1", "Trait2", "VarA"), 3) == -863
    nseg = {(r["sireID"], r["damID"]): r["Nsegsnps"] for r in res}
    assert nseg[("146", "146")] == 34 and nseg[("153", "146")] == 48
    # tidy BV rows :588-593
    assert sig(g("49", "49", "Trait1", "Trait1", "VarA"), 3) == 640
    assert sig(g("49", "49", "Trait2", "Trait2", "VarA"), 4) == 1597
    assert sig(g("49", "153", "Trait1", "Trait1", "VarA"), 4) == 1064
    assert sig(g("49", "153", "Trait2", "Trait2", "VarA"), 4) == 1815
    # SELIND BV / TGV (= BV + DD) :588,591,600-605
    SI = {"Trait1": 0.75, "Trait2": 0.25}
    want = {("49", "49"): (334, None), ("49", "153"): (479, None), ("146", "146"): (509, 521),
            ("153", "146"): (567, 579), ("153", "153"): (624, 640)}
    for r in res:
        key = (r["sireID"], r["damID"])
        if key not in want:
            continue
        bv = [(a, c, v) for a, c, p, v in r["predVars"] if p == "VarA"]
        dd = [(a, c, v) for a, c, p, v in r["predVars"] if p == "VarD"]
        tgv = [(a, c, v + w) for (a, c, v), (_, _, w) in zip(bv, dd)]
        assert sig(O.selind_var(bv, b.traits, SI), 3) == want[key][0]
        if want[key][1] is not None:
            assert sig(O.selind_var(tgv, b.traits, SI), 3) == want[key][1]


def test_single_trait_has_no_covariances(bundled):
    b = bundled
    res = run(b, "A", b.effdict(b.effA, ["Trait1"]))
    assert all(len(r["predVars"]) == 1 for r in res)             # R/predCrossVar.R:132
    assert sig(res[0]["predVars"][0][3], 3) == 707


def test_genmap2recombfreq_follows_current_source(bundled):
    b = bundled
    c1 = O.genmap2recombfreq(b.genmap, b.chrom)
    assert np.all(c1[np.ix_(b.chrom == 1, b.chrom == 2)] == 0.5)
    R = O.recombfreq_to_ld_decay(c1)
    assert np.all(R[np.ix_(b.chrom == 1, b.chrom == 2)] == 0.0)
    assert np.all(np.diag(R) == 1.0)
    j, k = 3, 17
    assert R[j, k] == pytest.approx(np.exp(-2 * abs(b.genmap[j] - b.genmap[k]) / 100), rel=1e-14)
    # fixture quirk: the bundled matrix was made WITHOUT the /100
    same = b.chrom[:, None] == b.chrom[None, :]
    old = np.where(same, np.exp(-2 * np.abs(b.genmap[:, None] - b.genmap[None, :])), 0.0)
    assert np.max(np.abs(old - b.recombFreqMat)) < 1e-15


def test_delta_identity_used_by_the_kernels(bundled):
    """0.5 X'X - pp' == 0.25 delta delta'  (exact)."""
    b = bundled
    for gid in b.PARENTS:
        i = b.hap_rownames.index(f"{gid}_HapA")
        k = b.hap_rownames.index(f"{gid}_HapB")
        X = b.haploMat[[i, k]].astype(np.float64)
        d = X[0] - X[1]
        assert np.array_equal(O.calcGameticLD(X, b.recombFreqMat), b.recombFreqMat * (0.25 * np.outer(d, d)))


def test_missing_parent_raises(bundled):
    b = bundled
    with pytest.raises(KeyError):
        O.predCrossVars([("49", "nope")], "A", b.effdict(b.effA), haploMat=b.haploMat,
                        hap_rownames=b.hap_rownames, recombFreqMat=b.recombFreqMat)
