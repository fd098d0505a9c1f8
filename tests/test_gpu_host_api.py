"""GPU tests of the host-side mirror of the R API (predCrossVars / predictCrosses / predictCrossVars /
predCrossMeans on pandas objects): name matching, output schema and row order, error behaviour,
and the printed vignette values (docs/articles/*.html) end to end."""
import numpy as np
import pandas as pd
import pytest

from oracle import predcrossvar_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import genomicmateselectr_b200 as g
    return g


@pytest.fixture(scope="module")
def objs(bundled):
    b = bundled
    hap = pd.DataFrame(b.haploMat.astype(np.int32), index=b.hap_rownames, columns=b.snp_names)
    R = pd.DataFrame(b.recombFreqMat, index=b.snp_names, columns=b.snp_names)
    dose = pd.DataFrame(b.doseMat.astype(np.int32), index=b.dose_rownames, columns=b.snp_names)
    eff = lambda mat: {t: pd.Series(mat[:, i], index=b.snp_names) for i, t in enumerate(b.traits)}
    return dict(hap=hap, R=R, dose=dose, eff=eff)


def sig(x, n):
    from math import floor, log10
    return 0.0 if x == 0 else round(x, n - 1 - floor(log10(abs(x))))


def test_predCrossVars_schema_order_and_values(g, bundled, objs):
    b = bundled
    crosses = g.crosses2predict(b.PARENTS)
    assert list(crosses.iloc[1]) == ["49", "153"] and len(crosses) == 15
    crosses["note"] = [f"row{i}" for i in range(15)]                     # extra columns are carried through
    res = g.predCrossVars(CrossesToPredict=crosses, modelType="AD", AddEffectList=objs["eff"](b.effAD_add),
                          DomEffectList=objs["eff"](b.effAD_dom), haploMat=objs["hap"], recombFreqMat=objs["R"],
                          ncores=4, nBLASthreads=2)
    assert list(res.columns) == ["sireID", "damID", "note", "Nsegsnps", "ComputeTime", "predVars"]
    assert len(res) == 15 and list(res.Nsegsnps[:4]) == [22, 46, 48, 53] and res.note[3] == "row3"
    pv = res.predVars[14]                                                # 146 x 146
    assert list(pv.columns) == ["Trait1", "Trait2", "predOf", "predVar"]
    assert list(zip(pv.Trait1, pv.Trait2, pv.predOf)) == [
        ("Trait1", "Trait1", "VarA"), ("Trait1", "Trait1", "VarD"), ("Trait2", "Trait2", "VarA"),
        ("Trait2", "Trait2", "VarD"), ("Trait1", "Trait2", "VarA"), ("Trait1", "Trait2", "VarD")]
    assert [sig(v, 3) for v in pv.predVar] == [1230, 9.97, 1980, 51.3, -820, 9.04]   # non_additive_models.html:624-629
    want = O.predCrossVars(list(zip(crosses.sireID, crosses.damID)), "AD", {t: v.to_numpy() for t, v in objs["eff"](b.effAD_add).items()},
                           {t: v.to_numpy() for t, v in objs["eff"](b.effAD_dom).items()}, haploMat=b.haploMat,
                           hap_rownames=b.hap_rownames, recombFreqMat=b.recombFreqMat)
    for row, w in zip(res.itertuples(), want):
        assert (row.sireID, row.damID, row.Nsegsnps) == (w["sireID"], w["damID"], w["Nsegsnps"])
        got = list(zip(row.predVars.Trait1, row.predVars.Trait2, row.predVars.predOf))
        assert got == [(a, c, p) for a, c, p, _ in w["predVars"]]
        assert np.allclose(row.predVars.predVar, [v for *_, v in w["predVars"]], rtol=1e-9, atol=0)


def test_names_drive_the_alignment(g, bundled, objs):
    b = bundled
    crosses = pd.DataFrame({"damID": ["153", "146"], "sireID": ["49", "146"]})      # column order is irrelevant
    base = g.predCrossVars(crosses, "A", objs["eff"](b.effA), haploMat=objs["hap"], recombFreqMat=objs["R"])
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(b.snp_names))
    eff = {t: s.iloc[rng.permutation(len(s))] for t, s in objs["eff"](b.effA).items()}   # effects in another SNP order
    hap = objs["hap"].iloc[:, perm]                                                       # haploMat columns shuffled
    R = objs["R"].iloc[rng.permutation(100), rng.permutation(100)]                       # matrix rows / cols shuffled
    shuf = g.predCrossVars(crosses, "A", eff, haploMat=hap, recombFreqMat=R)
    assert list(shuf.Nsegsnps) == list(base.Nsegsnps) == [46, 34]
    for a, c in zip(base.predVars, shuf.predVars):
        assert np.allclose(a.predVar, c.predVar, rtol=1e-11)
    assert sig(base.predVars[0].predVar[0], 4) == 1079                                  # start_here.html:385
    # modelType other than "AD" behaves as "A" (R/predCrossVar.R:33,39,119,216)
    odd = g.predCrossVars(crosses, "DirDom", objs["eff"](b.effA), haploMat=objs["hap"], recombFreqMat=objs["R"])
    assert np.array_equal(odd.predVars[0].predVar, base.predVars[0].predVar)


def test_errors_like_the_reference(g, bundled, objs):
    b = bundled
    with pytest.raises(KeyError):                                                      # subscript out of bounds, :46
        g.predCrossVars(pd.DataFrame({"sireID": ["49"], "damID": ["nobody"]}), "A", objs["eff"](b.effA),
                        haploMat=objs["hap"], recombFreqMat=objs["R"])
    with pytest.raises(ValueError, match="PMV"):
        g.predCrossVars(g.crosses2predict(["49"]), "A", objs["eff"](b.effA), predType="PMV",
                        haploMat=objs["hap"], recombFreqMat=objs["R"])
    eff = {t: s.iloc[:50] for t, s in objs["eff"](b.effA).items()}
    with pytest.raises(KeyError):
        g.predCrossVars(g.crosses2predict(["49"]), "A", eff, haploMat=objs["hap"], recombFreqMat=objs["R"])


def test_zero_segregation_cross_vanishes(g, bundled, objs):
    b = bundled
    hap = objs["hap"].copy()
    hap.loc[["49_HapA", "49_HapB"]] = 1
    res = g.predCrossVars(pd.DataFrame({"sireID": ["49", "49"], "damID": ["49", "153"]}), "A", objs["eff"](b.effA),
                          haploMat=hap, recombFreqMat=objs["R"])
    assert len(res) == 1 and res.damID[0] == "153"                                     # R/predCrossVar.R:156-160,70-71


def test_lazy_genmap_path(g, bundled, objs):
    b = bundled
    m = pd.Series(b.genmap, index=b.snp_names)
    rf = 1 - (2 * g.genmap2recombfreq(m, nChr=2))                                        # what users write (start_here.Rmd:112)
    assert isinstance(rf, g.LDDecay)
    crosses = g.crosses2predict(b.PARENTS[:3])
    res = g.predCrossVars(crosses, "AD", objs["eff"](b.effAD_add), objs["eff"](b.effAD_dom), haploMat=objs["hap"],
                          recombFreqMat=rf)
    Rcur = O.recombfreq_to_ld_decay(O.genmap2recombfreq(b.genmap, b.chrom))              # current R/helpers.R:204-213
    want = O.predCrossVars(list(zip(crosses.sireID, crosses.damID)), "AD", {t: b.effAD_add[:, i] for i, t in enumerate(b.traits)},
                           {t: b.effAD_dom[:, i] for i, t in enumerate(b.traits)}, haploMat=b.haploMat,
                           hap_rownames=b.hap_rownames, recombFreqMat=Rcur)
    for row, w in zip(res.itertuples(), want):
        assert np.allclose(row.predVars.predVar, [v for *_, v in w["predVars"]], rtol=1e-9, atol=0)


def snpeffs_frame(b, objs, model):
    d = {"Trait": b.traits}
    if model == "A":
        d["allelesubsnpeff"] = list(objs["eff"](b.effA).values())
    elif model == "AD":
        d["allelesubsnpeff"] = list(objs["eff"](b.effAD_add).values())
        d["domdevsnpeff"] = list(objs["eff"](b.effAD_dom).values())
    else:
        a, dd, alpha = b.dirdom_effects()
        d["allelesubsnpeff"] = list(objs["eff"](alpha).values())
        d["addsnpeff"] = list(objs["eff"](a).values())
        d["domsnpeff"] = list(objs["eff"](dd).values())
    return pd.DataFrame(d)


def test_predictCrosses_A_matches_vignette(g, bundled, objs):
    b = bundled
    out = g.predictCrosses(modelType="A", selInd=True, SIwts={"Trait1": 0.75, "Trait2": 0.25},
                           CrossesToPredict=g.crosses2predict(b.PARENTS), snpeffs=snpeffs_frame(b, objs, "A"),
                           dosages=objs["dose"], haploMat=objs["hap"], recombFreqMat=objs["R"])
    tidy = out["tidyPreds"]
    assert list(tidy.columns) == ["sireID", "damID", "Nsegsnps", "predOf", "Trait", "predMean", "predVar", "predSD", "predUsefulness"]
    assert len(tidy) == 45 and set(tidy.predOf) == {"BV"}
    first = tidy.iloc[:3]                                    # docs/articles/start_here.html:333-335
    assert list(first.Trait) == ["SELIND", "Trait1", "Trait2"]
    assert [sig(v, 3) for v in first.predMean] == [115, 194, -119]
    assert [sig(v, 3) for v in first.predVar] == [361, 707, 1520]
    assert [sig(v, 3) for v in first.predSD] == [19.0, 26.6, 39.0]
    assert [sig(v, 3) for v in first.predUsefulness] == [166, 265, -15.1]
    raw = out["rawPreds"]["predVars"]
    assert len(raw) == 45 and set(raw.predOf) == {"VarBV"}
    assert [sig(v, 3) for v in raw.predVar[:3]] == [707, 1520, -350]       # :385


def test_predictCrosses_AD_and_DirDom(g, bundled, objs):
    b = bundled
    crosses = g.crosses2predict(b.PARENTS)
    SI = {"Trait1": 0.75, "Trait2": 0.25}
    out = g.predictCrosses("AD", selInd=True, SIwts=SI, CrossesToPredict=crosses, snpeffs=snpeffs_frame(b, objs, "AD"),
                           dosages=objs["dose"], haploMat=objs["hap"], recombFreqMat=objs["R"])
    t = out["tidyPreds"]
    assert len(t) == 90 and set(t.predOf) == {"BV", "TGV"}
    q = t[(t.sireID == "146") & (t.damID == "146") & (t.Trait == "SELIND")].set_index("predOf")
    assert sig(q.predVar["BV"], 3) == 509 and sig(q.predVar["TGV"], 3) == 521      # non_additive_models.html:600-601
    assert sig(q.predMean["BV"], 3) == 82.1
    q = t[(t.sireID == "49") & (t.damID == "49") & (t.predOf == "BV")].set_index("Trait")
    assert [sig(q.predVar[k], 3) for k in ("SELIND", "Trait1", "Trait2")] == [334, 640, 1600]   # :588-590
    assert sig(q.predMean["Trait1"], 3) == 203 and sig(q.predUsefulness["SELIND"], 3) == 179
    dd = g.predictCrosses("DirDom", selInd=True, SIwts=SI, CrossesToPredict=crosses, snpeffs=snpeffs_frame(b, objs, "DirDom"),
                          dosages=objs["dose"], haploMat=objs["hap"], recombFreqMat=objs["R"])
    assert len(dd["tidyPreds"]) == 90                                                  # :660 tibble [90 x 9]
    raw = dd["rawPreds"]["predVars"]
    assert list(raw.predOf[:3]) == ["VarBV"] * 3 and set(raw.predOf) == {"VarBV", "VarA", "VarD"}
    assert len(raw) == 15 * 3 * 3
    means = dd["rawPreds"]["predMeans"]
    assert set(means.predOf) == {"MeanBV", "MeanTGV"}
    # MeanTGV against the oracle's Falconer-MacKay restatement
    a, d_, _ = b.dirdom_effects()
    i, k = b.dose_rownames.index("49"), b.dose_rownames.index("153")
    want = O.predCrossMeans_TGV(b.doseMat[i], b.doseMat[k], a[:, 0], d_[:, 0])
    got = means[(means.sireID == "49") & (means.damID == "153") & (means.Trait == "Trait1") & (means.predOf == "MeanTGV")]
    assert float(got.predMean.iloc[0]) == pytest.approx(want, rel=1e-10)


def test_predictCrossVars_cv_entry(g, bundled, objs):
    b = bundled
    eff = pd.DataFrame({"Dataset": ["trainset"] * 2 + ["testset"] * 2, "Trait": b.traits * 2,
                        "allelesubsnpeff": list(objs["eff"](b.effAD_add).values()) * 2,
                        "domdevsnpeff": list(objs["eff"](b.effAD_dom).values()) * 2})
    snpeffs = pd.DataFrame({"Repeat": ["Repeat1", "Repeat1"], "Fold": ["Fold1", "Fold2"], "effects": [eff, eff]})
    folds = pd.DataFrame({"Repeat": ["Repeat1", "Repeat1"], "Fold": ["Fold1", "Fold2"],
                          "CrossesToPredict": [g.crosses2predict(["49", "153"])[["damID", "sireID"]],
                                               g.crosses2predict(["74", "146"])]})
    res = g.predictCrossVars("AD", snpeffs, folds, objs["hap"], objs["R"], ncores=1)
    assert list(res.Fold) == ["Fold1", "Fold2"] and len(res.predVars[0]) == 3
    pv = res.predVars[1]
    last = pv[(pv.sireID == "146") & (pv.damID == "146")].predVars.iloc[0]
    assert set(last.predOf) == {"VarBV", "VarDD"}
    assert sig(float(last[(last.Trait1 == "Trait1") & (last.Trait2 == "Trait1") & (last.predOf == "VarBV")].predVar.iloc[0]), 4) == 1232
