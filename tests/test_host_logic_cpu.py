"""The host-side mirror of the R API (genomicmateselectr_b200/predcrossvar.py: name matching, SNP / chromosome ordering,
output schema and row order, error behaviour, the pandas tidy paths, CV batching) without a GPU: the device engine behind
CrossVarSession is replaced by a stand-in that answers from the oracle (tests/ may use oracle/; the product never does), and
the SAME test bodies as tests/test_gpu_host_api.py run against it - bundled data, printed vignette values.
What this does not cover is the engine itself: that is what the -m gpu files are for."""
import numpy as np
import pandas as pd
import pytest

from oracle import predcrossvar_oracle as O
from tests import test_gpu_host_api as G
from tests.util import oracle_slab, pair_index


def pair_slot(T):
    """(t, s), t <= s -> position of the pair in a slab row (variances first, then combn order)."""
    return {p: i for i, p in enumerate(pair_index(T))}


class OracleEngine:
    """Same methods and array conventions as genomicmateselectr_b200.engine.CrossVarEngine, evaluated by the NumPy oracle."""

    def __init__(self, device=0):
        self.R = self.H = None
        self.last = None
        self.closed = False

    # ---- geometry
    def set_genmap(self, cM, m_c):
        chrom = np.repeat(np.arange(1, len(m_c) + 1), np.asarray(m_c, dtype=np.int64))
        self.R = O.recombfreq_to_ld_decay(O.genmap2recombfreq(np.asarray(cM, dtype=np.float64), chrom))   # R/helpers.R:204-213

    def set_recomb_matrix(self, R, chrom):
        self.R = np.asarray(R, dtype=np.float64)
        chrom = np.asarray(chrom)
        off = chrom[:, None] != chrom[None, :]
        return not np.any(self.R[off] != 0.0)

    def set_haplotypes(self, H):
        H = np.asarray(H)
        if not np.all((H == 0) | (H == 1)):
            raise ValueError("haplotype values must be 0 / 1")
        self.H = H.astype(np.uint8)

    # ---- the path
    def predict(self, model, add, dom, asub, sire, dam):
        out, nseg = oracle_slab(model, self.H, self.R, add, dom, asub, sire, dam)
        self.last = dict(model=model, add=add, dom=dom, asub=asub, sire=np.asarray(sire), dam=np.asarray(dam), out=out, nseg=nseg)
        return out, nseg

    def predict_batched(self, model, add, dom, asub, sire, dam):
        outs = []
        for s in range(add.shape[0]):
            o, nseg = self.predict(model, add[s], None if dom is None else dom[s], None if asub is None else asub[s], sire, dam)
            outs.append(o)
        return np.stack(outs), nseg

    def cross_means(self, add, dom, sire, dam):
        """gebv[P, T], mean_bv[N, T], mean_tgv[N, T] | None (R/predCrossVar.R:336-399; dosage = HapA + HapB)."""
        dose = (self.H[0::2].astype(np.float64) + self.H[1::2])
        gebv = dose @ add
        mbv = 0.5 * (gebv[sire] + gebv[dam])
        mtgv = None
        if dom is not None:
            mtgv = np.array([[O.predCrossMeans_TGV(dose[s], dose[d], add[:, t], dom[:, t]) for t in range(add.shape[1])]
                             for s, d in zip(sire, dam)])
        return gebv, mbv, mtgv

    def tidy(self, w, std_sel_int):
        """tidy[x][predOf][slot][(mean, var, sd, usefulness)] as include/gms_b200.h documents gms_tidy
        (R/gmsFunctions.R:726-871), from the last predict()."""
        L = self.last
        out, model = L["out"], L["model"]
        T = L["add"].shape[1]
        if model == "A":
            var, bv_eff = [out[:, :, 0]], L["add"]
        elif model == "AD":
            var, bv_eff = [out[:, :, 0], out[:, :, 0] + out[:, :, 1]], L["add"]
        else:
            var, bv_eff = [out[:, :, 2], out[:, :, 0] + out[:, :, 1]], L["asub"]
        _, mbv, _ = self.cross_means(bv_eff, None, L["sire"], L["dam"])
        means = [mbv]
        if model == "AD":
            means.append(mbv)                                              # :737-741
        elif model == "DirDom":
            means.append(self.cross_means(L["add"], L["dom"], L["sire"], L["dam"])[2])
        nslot = T + (w is not None)
        tab = np.zeros((out.shape[0], len(var), nslot, 4))
        for oi, (v, mu) in enumerate(zip(var, means)):
            cols_m, cols_v = [], []
            if w is not None:                                              # w' G w and w' mean (:805-853)
                Gm = np.zeros((out.shape[0], T, T))
                slot = pair_slot(T)
                for t in range(T):
                    for s in range(t, T):
                        Gm[:, t, s] = Gm[:, s, t] = v[:, slot[(t, s)]]
                cols_v.append(np.einsum("t,xts,s->x", w, Gm, w))
                cols_m.append(mu @ w)
            for t in range(T):
                cols_v.append(v[:, t])
                cols_m.append(mu[:, t])
            vv, mm = np.stack(cols_v, axis=1), np.stack(cols_m, axis=1)
            sd = np.sqrt(vv)
            tab[:, oi] = np.stack([mm, vv, sd, mm + std_sel_int * sd], axis=2)
        return tab

    def close(self):
        self.closed = True


@pytest.fixture()
def g(monkeypatch):
    import genomicmateselectr_b200 as pkg
    from genomicmateselectr_b200 import predcrossvar
    monkeypatch.setattr(predcrossvar, "CrossVarEngine", OracleEngine)
    return pkg


@pytest.fixture(scope="module")
def objs(bundled):
    b = bundled
    hap = pd.DataFrame(b.haploMat.astype(np.int32), index=b.hap_rownames, columns=b.snp_names)
    R = pd.DataFrame(b.recombFreqMat, index=b.snp_names, columns=b.snp_names)
    dose = pd.DataFrame(b.doseMat.astype(np.int32), index=b.dose_rownames, columns=b.snp_names)
    eff = lambda mat: {t: pd.Series(mat[:, i], index=b.snp_names) for i, t in enumerate(b.traits)}
    return dict(hap=hap, R=R, dose=dose, eff=eff)


def test_pair_index_helper_matches_the_library_order():
    from genomicmateselectr_b200.engine import trait_pairs
    p1, p2 = trait_pairs(4)
    idx = pair_slot(4)
    assert all(idx[(int(a), int(b))] == i for i, (a, b) in enumerate(zip(p1, p2))) and len(idx) == len(p1) == 10


@pytest.mark.parametrize("body", [G.test_predCrossVars_schema_order_and_values, G.test_names_drive_the_alignment,
                                  G.test_errors_like_the_reference, G.test_zero_segregation_cross_vanishes,
                                  G.test_lazy_genmap_path, G.test_predictCrosses_A_matches_vignette,
                                  G.test_predictCrosses_AD_and_DirDom, G.test_predictCrossVars_cv_entry],
                         ids=lambda f: f.__name__)
def test_host_api_on_the_oracle_engine(body, g, bundled, objs):
    body(g, bundled, objs)


def test_pandas_tidy_paths_without_the_fused_table(g, bundled, objs):
    """predTheMeans / predTheVars alone take the host-side pandas route of R/gmsFunctions.R:764-871 (no gms_tidy): the
    selection-index variances and means must equal the vignette's (start_here.html:333-335)."""
    b = bundled
    kw = dict(modelType="A", selInd=True, SIwts={"Trait1": 0.75, "Trait2": 0.25}, CrossesToPredict=g.crosses2predict(b.PARENTS),
              snpeffs=G.snpeffs_frame(b, objs, "A"), dosages=objs["dose"], haploMat=objs["hap"], recombFreqMat=objs["R"])
    v = g.predictCrosses(predTheMeans=False, **kw)["tidyPreds"]
    assert list(v.columns) == ["sireID", "damID", "Nsegsnps", "predOf", "Trait", "predVar", "predSD"]
    assert list(v.Trait[:3]) == ["SELIND", "Trait1", "Trait2"] and [G.sig(x, 3) for x in v.predVar[:3]] == [361, 707, 1520]
    mns = g.predictCrosses(predTheVars=False, **kw)["tidyPreds"]
    assert list(mns.columns) == ["sireID", "damID", "predOf", "Trait", "predMean"]
    assert list(mns.Trait[:3]) == ["SELIND", "Trait1", "Trait2"] and [G.sig(x, 3) for x in mns.predMean[:3]] == [115, 194, -119]
    full = g.predictCrosses(**kw)["tidyPreds"]
    assert np.allclose(full.predVar, v.predVar, rtol=1e-12) and np.allclose(full.predMean, mns.predMean, rtol=1e-12)


def test_long_form_equals_the_unnested_nested_form(g, bundled, objs):
    """nest=False (one frame, what large cross lists use) is tidyr::unnest(predVars) of the reference's nested result,
    including the dropped zero-segregation cross and the carried extra column."""
    from genomicmateselectr_b200.predcrossvar import _unnest
    b = bundled
    hap = objs["hap"].copy()
    hap.loc[["49_HapA", "49_HapB"]] = 1                                  # 49 x 49 has no segregating SNP
    crosses = g.crosses2predict(b.PARENTS[:4])
    crosses["note"] = [f"row{i}" for i in range(len(crosses))]
    kw = dict(CrossesToPredict=crosses, modelType="AD", AddEffectList=objs["eff"](b.effAD_add),
              DomEffectList=objs["eff"](b.effAD_dom), haploMat=hap, recombFreqMat=objs["R"])
    nested, long = g.predCrossVars(**kw), g.predCrossVars(nest=False, **kw)
    assert len(nested) == len(crosses) - 1 and len(long) == len(nested) * 6
    flat = _unnest(nested)
    cols = ["sireID", "damID", "note", "Nsegsnps", "Trait1", "Trait2", "predOf"]
    assert list(long.columns) == list(flat.columns)
    assert flat[cols].astype(str).equals(long[cols].astype(str)) and np.array_equal(flat.predVar, long.predVar)
