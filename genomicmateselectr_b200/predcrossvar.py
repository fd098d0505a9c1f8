"""Host-side mirror of genomicMateSelectR's cross-variance API (same names, argument meaning,
output schema and error behaviour as the R functions), driving libgms_b200.so.

R is not installed in this image, so next to the R glue in r-package/ (which cannot be executed
here) this module is the executable host layer: pandas DataFrames stand in for tibbles and named
matrices.  No arithmetic of the hot path happens here - only name matching, index building and
tidy-table assembly; the numbers come from the CUDA library (no CPU fallback).

Reference: R/predCrossVar.R:24-78 (predCrossVars), R/gmsFunctions.R:650-880 (predictCrosses),
R/parentWiseCrossVal.R:386-493 (predictCrossVars), R/predCrossVar.R:336-399 (predCrossMeans),
R/helpers.R:204-213 (genmap2recombfreq), :278-289 (crosses2predict).
"""
from __future__ import annotations

import time

import numpy as np
import pandas as pd

from .engine import CrossVarEngine, trait_pairs

__all__ = ["RecombFreq", "LDDecay", "genmap2recombfreq", "crosses2predict", "CrossVarSession",
           "predCrossVars", "predCrossMeans", "predictCrosses", "predictCrossVars"]


# --------------------------------------------------------------------------- stage (1), lazily
class RecombFreq:
    """Result of genmap2recombfreq(): the pairwise recombination-frequency matrix c1, kept as the
    map itself.  `1 - (2 * c1)` (what the reference's users write, vignettes/start_here.Rmd:112)
    gives an LDDecay that predCrossVars hands to the device, which builds the per-chromosome
    blocks in HBM; the m x m matrix (8.7 GB at 33k SNPs) never exists on the host."""

    def __init__(self, cM: pd.Series, nChr=None):
        self.cM = cM
        self.nChr = nChr

    def __rmul__(self, k):
        if k != 2:
            raise ValueError("only 1-(2*c1) is supported on a lazy recombination-frequency matrix")
        return _TwoC(self)

    __mul__ = __rmul__


class _TwoC:
    def __init__(self, rf):
        self.rf = rf

    def __rsub__(self, one):
        if one != 1:
            raise ValueError("only 1-(2*c1) is supported on a lazy recombination-frequency matrix")
        return LDDecay(self.rf.cM)


class LDDecay:
    """recombFreqMat = 1-2c in map form (see RecombFreq)."""

    def __init__(self, cM: pd.Series):
        self.cM = cM

    @property
    def index(self):
        return self.cM.index


def genmap2recombfreq(m, nChr=None):
    """R/helpers.R:204-213.  `m`: Series of cM positions named "<chr>_<id>" (chr integer)."""
    if not isinstance(m, pd.Series):
        raise TypeError("m must be a pandas Series named by SNP id ('<chr>_<id>')")
    return RecombFreq(m.astype(np.float64), nChr)


def crosses2predict(parents):
    """R/helpers.R:278-289: all pairwise matings incl. selfs, upper triangle, sire-major."""
    parents = [str(p) for p in parents]
    rows = [(parents[i], parents[j]) for i in range(len(parents)) for j in range(i, len(parents))]
    return pd.DataFrame(rows, columns=["sireID", "damID"])


def _chrom_of(names):
    """Integer prefix before the first '_' (R/helpers.R:199,208-211); one block if unparsable."""
    out = []
    for n in names:
        head = str(n).split("_", 1)[0]
        if "_" not in str(n) or not head.lstrip("-").isdigit():
            return np.zeros(len(names), dtype=np.int64)
        out.append(int(head))
    return np.asarray(out, dtype=np.int64)


def _effect_matrix(EffectList, snp_names):
    """dict trait -> Series (index = SNP ids) or 1 x m DataFrame (columns = SNP ids) -> (traits, m x T),
    aligned to `snp_names` BY NAME (R/predCrossVar.R:117-120)."""
    traits = list(EffectList.keys())
    cols = []
    for t in traits:
        e = EffectList[t]
        if isinstance(e, pd.DataFrame):
            if e.shape[0] != 1:
                raise ValueError("VPM effect matrices must be 1 x m (predType='PMV' is not supported)")
            e = e.iloc[0]
        if not isinstance(e, pd.Series):
            raise TypeError("effects must be named: a Series indexed by SNP id or a 1 x m DataFrame")
        try:
            cols.append(e.loc[snp_names].to_numpy(dtype=np.float64))
        except KeyError as err:
            raise KeyError(f"effects of trait {t!r} lack SNPs present in haploMat: {err}") from None
    return traits, np.stack(cols, axis=1)


class CrossVarSession:
    """Device-resident state for many predCrossVars calls on the same haploMat / recombFreqMat
    (the cross-validation driver calls the path nrepeats x nfolds x (1-2) times).
    The R glue keeps the same thing behind an external pointer."""

    def __init__(self, haploMat: pd.DataFrame, recombFreqMat, parents=None, device=0):
        if not isinstance(haploMat, pd.DataFrame):
            raise TypeError("haploMat must be a DataFrame (index '<GID>_HapA'/'<GID>_HapB', columns = SNP ids)")
        snps = [str(s) for s in haploMat.columns]
        chrom = _chrom_of(snps)
        # group SNPs by chromosome, keeping first-appearance order of chromosomes and SNP order within
        first = {}
        for c in chrom:
            first.setdefault(int(c), len(first))
        order = np.argsort(np.array([first[int(c)] for c in chrom]), kind="stable")
        self.snp_names = [snps[i] for i in order]
        self.chrom = chrom[order]
        m_c = np.array([np.sum(self.chrom == c) for c in first], dtype=np.int64)
        self.engine = CrossVarEngine(device)
        if isinstance(recombFreqMat, LDDecay):
            try:
                cM = recombFreqMat.cM.loc[self.snp_names].to_numpy(np.float64)
            except KeyError as err:
                raise KeyError(f"genetic map lacks SNPs present in haploMat: {err}") from None
            self.engine.set_genmap(cM, m_c)
            self.block_diagonal = True
        else:
            if isinstance(recombFreqMat, pd.DataFrame):
                try:
                    R = recombFreqMat.loc[self.snp_names, self.snp_names].to_numpy(np.float64)
                except KeyError as err:
                    raise KeyError(f"recombFreqMat lacks SNPs present in haploMat: {err}") from None
            else:
                R = np.asarray(recombFreqMat, dtype=np.float64)[np.ix_(order, order)]
            self.block_diagonal = self.engine.set_recomb_matrix(R, self.chrom)
        # haplotypes of the requested parents (default: every GID that has both rows)
        rows = {str(r): i for i, r in enumerate(haploMat.index)}
        if parents is None:
            parents = [r[:-5] for r in rows if r.endswith("_HapA") and r[:-5] + "_HapB" in rows]
        self.parents = [str(p) for p in parents]
        idx = []
        for p in self.parents:
            try:
                idx += [rows[f"{p}_HapA"], rows[f"{p}_HapB"]]     # exact match, cf. R/predCrossVar.R:262
            except KeyError as err:
                raise KeyError(f"parent {p!r}: haplotype row {err} not in haploMat") from None
        H = haploMat.to_numpy()[np.asarray(idx, dtype=np.int64)][:, order]
        self.engine.set_haplotypes(H)
        self.parent_index = {p: i for i, p in enumerate(self.parents)}

    def indices(self, ids):
        try:
            return np.array([self.parent_index[str(i)] for i in ids], dtype=np.int32)
        except KeyError as err:
            raise KeyError(f"parent {err} is not in haploMat (subscript out of bounds)") from None

    def close(self):
        self.engine.close()


def _run(session, CrossesToPredict, model, AddEffectList, DomEffectList=None, AlleleSubEffectList=None):
    traits, add = _effect_matrix(AddEffectList, session.snp_names)
    dom = asub = None
    if model in ("AD", "DirDom"):
        t2, dom = _effect_matrix(DomEffectList, session.snp_names)
        if t2 != traits:
            raise ValueError("AddEffectList and DomEffectList must name the same traits in the same order")
    if model == "DirDom":
        t3, asub = _effect_matrix(AlleleSubEffectList, session.snp_names)
        if t3 != traits:
            raise ValueError("AlleleSubEffectList must name the same traits in the same order")
    sire = session.indices(CrossesToPredict["sireID"])
    dam = session.indices(CrossesToPredict["damID"])
    out, nseg = session.engine.predict(model, add, dom, asub, sire, dam)
    return traits, out, nseg


def _pair_labels(traits, comp_names):
    """(Trait1, Trait2, predOf) of one cross's rows in the reference's order: T variances then combn(traits, 2), the
    components of a pair adjacent (R/predCrossVar.R:128-151, :231-246)."""
    p1, p2 = trait_pairs(len(traits))
    t1 = np.array([traits[i] for i in p1 for _ in comp_names], dtype=object)
    t2 = np.array([traits[i] for i in p2 for _ in comp_names], dtype=object)
    of = np.array([c for _ in p1 for c in comp_names], dtype=object)
    return t1, t2, of


def _long(CrossesToPredict, traits, out, nseg, comp_names, elapsed):
    """tidyr::unnest(predVars) of the reference's result, built in one piece: one row per (cross with >= 1 segregating
    SNP, trait pair, component). No per-cross objects, so it scales to millions of crosses."""
    t1, t2, of = _pair_labels(traits, comp_names)
    k = len(t1)
    keep = np.flatnonzero(nseg > 0)
    res = CrossesToPredict.iloc[np.repeat(keep, k)].reset_index(drop=True).copy()
    res["Nsegsnps"] = np.repeat(nseg[keep].astype(np.int64), k)
    res["ComputeTime"] = elapsed / max(1, len(CrossesToPredict))
    res["Trait1"] = np.tile(t1, len(keep))
    res["Trait2"] = np.tile(t2, len(keep))
    res["predOf"] = np.tile(of, len(keep))
    res["predVar"] = out[keep].reshape(-1)
    return res


def _tidy(CrossesToPredict, traits, out, nseg, comp_names, elapsed):
    """One row per cross with >= 1 segregating SNP (R/predCrossVar.R:152-160,70-71); predVars holds
    (Trait1, Trait2, predOf, predVar) in the reference's row order (:128-151, :231-246)."""
    t1, t2, of = _pair_labels(traits, comp_names)
    keep = np.flatnonzero(nseg > 0)
    res = CrossesToPredict.iloc[keep].reset_index(drop=True).copy()
    res["Nsegsnps"] = nseg[keep].astype(np.int64)
    res["ComputeTime"] = elapsed / max(1, len(CrossesToPredict))
    res["predVars"] = [pd.DataFrame({"Trait1": t1, "Trait2": t2, "predOf": of, "predVar": out[x].reshape(-1)})
                       for x in keep]
    return res


def predCrossVars(CrossesToPredict, modelType, AddEffectList, DomEffectList=None, predType="VPM",
                  haploMat=None, recombFreqMat=None, ncores=1, nBLASthreads=None, session=None, nest=True, **kwargs):
    """Drop-in for R/predCrossVar.R:24-78.  Same arguments; `ncores` / `nBLASthreads` are accepted
    and ignored (the crosses are batched on the GPU); `session=` optionally reuses device state.
    Returns a DataFrame: sireID, damID, [extra columns], Nsegsnps, ComputeTime, predVars (one small frame per cross,
    the reference's nested schema). `nest=False` returns the unnested long form (Trait1, Trait2, predOf, predVar as
    columns) in one piece - use it for large cross lists, where a Python object per cross would dominate."""
    start = time.perf_counter()
    if predType != "VPM":
        raise ValueError("predType='PMV' is not supported (it is broken in the reference as well: "
                         "R/predCrossVar.R:189-202 never receives segsnps2keep)")
    if not isinstance(CrossesToPredict, pd.DataFrame) or not {"sireID", "damID"} <= set(CrossesToPredict.columns):
        raise ValueError("CrossesToPredict needs columns sireID and damID")
    model = "AD" if modelType == "AD" else "A"          # anything else behaves as "A" (R/predCrossVar.R:33,39,119,216)
    own = session is None
    if own:
        parents = pd.unique(pd.concat([CrossesToPredict["sireID"], CrossesToPredict["damID"]]).astype(str))  # :44-46
        session = CrossVarSession(haploMat, recombFreqMat, parents=parents)
    try:
        traits, out, nseg = _run(session, CrossesToPredict, model, AddEffectList,
                                 DomEffectList if model == "AD" else None)
    finally:
        if own:
            session.close()
    elapsed = time.perf_counter() - start
    comps = ["VarA", "VarD"][: 2 if model == "AD" else 1]
    res = (_tidy if nest else _long)(CrossesToPredict, traits, out, nseg, comps, elapsed)
    print(f"Done predicting fam vars. Took {round(elapsed / 60, 2)} mins for {int(np.sum(nseg > 0))} crosses")   # :74-76
    return res


def _unnest(df):
    """tidyr::unnest(predVars)."""
    parts = []
    base = df.drop(columns=["predVars"])
    for i, pv in enumerate(df["predVars"]):
        rep = base.iloc[[i] * len(pv)].reset_index(drop=True)
        parts.append(pd.concat([rep, pv.reset_index(drop=True)], axis=1))
    if not parts:
        return base.assign(Trait1=[], Trait2=[], predOf=[], predVar=[])
    return pd.concat(parts, ignore_index=True)


def predCrossMeans(CrossesToPredict, predType, AddEffectList, DomEffectList=None, doseMat=None, ncores=1,
                   session=None, **kwargs):
    """R/predCrossVar.R:336-399 on the device-resident haplotypes (dosage = HapA + HapB; `doseMat`
    is accepted for signature compatibility and only used to check that the parents exist).
    predType "BV": sireGEBV, damGEBV, predMean = mid-parent; "TGV": Falconer-MacKay eqn 14.6."""
    if session is None:
        raise ValueError("predCrossMeans needs session= (device-resident haplotypes)")
    traits, add = _effect_matrix(AddEffectList, session.snp_names)
    sire = session.indices(CrossesToPredict["sireID"])
    dam = session.indices(CrossesToPredict["damID"])
    if doseMat is not None:
        missing = set(map(str, CrossesToPredict["sireID"])) | set(map(str, CrossesToPredict["damID"]))
        missing -= set(map(str, doseMat.index))
        if missing:
            raise KeyError(f"parents not in doseMat: {sorted(missing)}")
    rows = []
    if predType == "BV":
        gebv, mbv, _ = session.engine.cross_means(add, None, sire, dam)
        for ti, t in enumerate(traits):
            d = CrossesToPredict.reset_index(drop=True).copy()
            d.insert(0, "Trait", t)
            d["sireGEBV"] = gebv[sire, ti]
            d["damGEBV"] = gebv[dam, ti]
            d["predOf"] = "MeanBV"
            d["predMean"] = mbv[:, ti]
            rows.append(d)
    elif predType == "TGV":
        _, dom = _effect_matrix(DomEffectList, session.snp_names)
        _, _, mtgv = session.engine.cross_means(add, dom, sire, dam)
        for ti, t in enumerate(traits):
            d = CrossesToPredict.reset_index(drop=True).copy()
            d.insert(0, "Trait", t)
            d["predOf"] = "MeanTGV"
            d["predMean"] = mtgv[:, ti]
            rows.append(d)
    else:
        raise ValueError("predType must be 'BV' or 'TGV'")
    return pd.concat(rows, ignore_index=True)


def _effect_lists(snpeffs: pd.DataFrame, col):
    """snpeffs$<col> -> named list of 1 x m matrices (R/gmsFunctions.R:661-678)."""
    if col not in snpeffs.columns:
        raise KeyError(f"snpeffs lacks column {col!r}")
    return {str(t): e for t, e in zip(snpeffs["Trait"], snpeffs[col])}


def predictCrosses(modelType, stdSelInt=2.67, selInd=False, SIwts=None, CrossesToPredict=None, snpeffs=None,
                   dosages=None, haploMat=None, recombFreqMat=None, ncores=1, nBLASthreads=None,
                   predTheMeans=True, predTheVars=True, session=None):
    """R/gmsFunctions.R:650-880.  `snpeffs`: DataFrame with column Trait and list-columns
    allelesubsnpeff [, domdevsnpeff | addsnpeff + domsnpeff] of Series indexed by SNP id.
    Returns {"tidyPreds": DataFrame, "rawPreds": {"predMeans": .., "predVars": ..}}.
    DirDom runs ONE device pass (3 components) instead of the reference's two predCrossVars calls."""
    if modelType not in ("A", "AD", "DirDom"):
        raise ValueError("modelType must be 'A', 'AD' or 'DirDom'")
    own = session is None
    if own:
        parents = pd.unique(pd.concat([CrossesToPredict["sireID"], CrossesToPredict["damID"]]).astype(str))
        session = CrossVarSession(haploMat, recombFreqMat, parents=parents)
    try:
        asub = _effect_lists(snpeffs, "allelesubsnpeff")
        raw = {}
        predictedvars = predictedmeans = fused = None
        if predTheVars:
            print("Predicting cross variance parameters")
            start = time.perf_counter()
            if modelType == "A":
                traits, out, nseg = _run(session, CrossesToPredict, "A", asub)
                comps = ["VarBV"]
            elif modelType == "AD":
                traits, out, nseg = _run(session, CrossesToPredict, "AD", asub, _effect_lists(snpeffs, "domdevsnpeff"))
                comps = ["VarBV", "VarDD"]
            else:
                traits, out, nseg = _run(session, CrossesToPredict, "DirDom", _effect_lists(snpeffs, "addsnpeff"),
                                         _effect_lists(snpeffs, "domsnpeff"), asub)
                # reference row order: all VarBV rows of a cross first, then VarA/VarD (R/gmsFunctions.R:719-723)
                comps = ["VarA", "VarD", "VarBV"]
            elapsed = time.perf_counter() - start
            print(f"Done predicting fam vars. Took {round(elapsed / 60, 2)} mins for {int(np.sum(nseg > 0))} crosses")
            predictedvars = _long(CrossesToPredict, traits, out, nseg, comps, elapsed)
            if predTheMeans:
                # the tidy table (TGV, SELIND, predSD, usefulness) fused on the device from the slab that is still in HBM
                w = None if not selInd else pd.Series(SIwts, dtype=np.float64).reindex(traits).fillna(0.0).to_numpy()
                fused = (session.engine.tidy(w, stdSelInt), nseg, traits)
            if modelType == "DirDom":
                bv = predictedvars[predictedvars.predOf == "VarBV"]
                predictedvars = pd.concat([bv, predictedvars[predictedvars.predOf != "VarBV"]], ignore_index=True)
            raw["predVars"] = predictedvars.copy()
        if predTheMeans:
            print("Predicting cross means")
            predictedmeans = predCrossMeans(CrossesToPredict, "BV", asub, doseMat=dosages, session=session)
            if modelType == "AD":
                predictedmeans = pd.concat([predictedmeans, predictedmeans.assign(predOf="TGV")], ignore_index=True)
            if modelType == "DirDom":
                predictedmeans = pd.concat([predictedmeans, predCrossMeans(
                    CrossesToPredict, "TGV", _effect_lists(snpeffs, "addsnpeff"),
                    _effect_lists(snpeffs, "domsnpeff"), doseMat=dosages, session=session)], ignore_index=True)
            raw["predMeans"] = predictedmeans.copy()
    finally:
        if own:
            session.close()

    # ---- tidy (R/gmsFunctions.R:764-871)
    if predTheMeans and predTheVars:
        tab, nseg, traits = fused            # [N, predOf, slot, (mean, var, sd, usefulness)] from gms_tidy
        keep = np.flatnonzero(nseg > 0)
        ofs = ["BV"] if modelType == "A" else ["BV", "TGV"]
        slots = (["SELIND"] if selInd else []) + list(traits)
        base = CrossesToPredict.iloc[np.repeat(keep, len(slots))][["sireID", "damID"]].reset_index(drop=True)
        parts = []
        for oi, of in enumerate(ofs):         # the reference binds all BV rows, then all TGV rows (:777-799)
            blk = base.copy()
            blk["Nsegsnps"] = np.repeat(nseg[keep].astype(np.int64), len(slots))
            blk["predOf"] = of
            blk["Trait"] = np.tile(np.array(slots, dtype=object), len(keep))
            v = tab[keep, oi].reshape(-1, 4)
            blk["predMean"], blk["predVar"], blk["predSD"], blk["predUsefulness"] = v[:, 0], v[:, 1], v[:, 2], v[:, 3]
            parts.append(blk)
        return {"tidyPreds": pd.concat(parts, ignore_index=True), "rawPreds": raw}
    if predTheMeans:
        pm = predictedmeans.copy()
        pm["predOf"] = pm["predOf"].str.replace("Mean", "", regex=False)
        pm = pm.rename(columns={"Trait": "Trait1"})
        pm["Trait2"] = pm["Trait1"]
        pm = pm[["sireID", "damID", "predOf", "Trait1", "Trait2", "predMean"]]
    if predTheVars:
        pv = predictedvars[["sireID", "damID", "Nsegsnps", "predOf", "Trait1", "Trait2", "predVar"]].copy()
        pv["predOf"] = pv["predOf"].str.replace("Var", "", regex=False)
        key = ["sireID", "damID", "Nsegsnps", "Trait1", "Trait2"]
        if modelType == "AD":
            bv, dd = pv[pv.predOf == "BV"], pv[pv.predOf == "DD"]
            tgv = bv.merge(dd, on=key, suffixes=("BV", "DD"))
            tgv = tgv.assign(predVar=tgv.predVarBV + tgv.predVarDD, predOf="TGV")[key[:3] + ["predOf"] + key[3:] + ["predVar"]]
            pv = pd.concat([bv, tgv], ignore_index=True)
        if modelType == "DirDom":
            bv, a, d = pv[pv.predOf == "BV"], pv[pv.predOf == "A"], pv[pv.predOf == "D"]
            tgv = a.merge(d, on=key, suffixes=("A", "D"))
            tgv = tgv.assign(predVar=tgv.predVarA + tgv.predVarD, predOf="TGV")[key[:3] + ["predOf"] + key[3:] + ["predVar"]]
            pv = pd.concat([bv, tgv], ignore_index=True)
    if selInd:
        print("Computing SELECTION INDEX means and variances.")
        w = pd.Series(SIwts, dtype=np.float64)
        if predTheMeans:
            # spread -> SELIND column after predOf -> pivot_longer (R/gmsFunctions.R:808-822): per (cross, predOf) the rows are
            # SELIND, then the traits. (Keys stay in cross-list order here; tidyr::spread sorts them.)
            mtraits = [str(t) for t in pd.unique(pm["Trait1"])]
            wide = pm.pivot_table(index=["sireID", "damID", "predOf"], columns="Trait1", values="predMean", sort=False)[mtraits]
            si = (wide[list(w.index)] * w.values).sum(axis=1).to_numpy()
            vals = np.column_stack([si, wide.to_numpy()])                       # [key, (SELIND, traits...)]
            keys = wide.index.to_frame(index=False)
            pm = keys.iloc[np.repeat(np.arange(len(keys)), 1 + len(mtraits))].reset_index(drop=True)
            pm["Trait1"] = np.tile(np.array(["SELIND"] + mtraits, dtype=object), len(keys))
            pm["Trait2"] = pm["Trait1"]
            pm["predMean"] = vals.reshape(-1)
        if predTheVars:
            parts = []
            for (s, d, n, of), grp in pv.groupby(["sireID", "damID", "Nsegsnps", "predOf"], sort=False):
                G = pd.DataFrame(np.nan, index=w.index, columns=w.index)
                for a, b, v in zip(grp.Trait1, grp.Trait2, grp.predVar):
                    if a in G.index and b in G.index:
                        G.loc[a, b] = v
                        G.loc[b, a] = v                                   # gmat[lower.tri] <- t(gmat)[lower.tri]
                v = float(w.values @ G.to_numpy() @ w.values)
                parts.append(pd.DataFrame({"sireID": [s], "damID": [d], "Nsegsnps": [n], "predOf": [of],
                                           "Trait1": ["SELIND"], "Trait2": ["SELIND"], "predVar": [v]}))
                parts.append(grp)
            pv = pd.concat(parts, ignore_index=True)
    if predTheMeans and predTheVars:
        tidy = pv[pv.Trait1 == pv.Trait2].merge(pm, on=["sireID", "damID", "predOf", "Trait1", "Trait2"])
        tidy = tidy.rename(columns={"Trait1": "Trait"})[["sireID", "damID", "Nsegsnps", "predOf", "Trait", "predMean", "predVar"]]
        tidy["predSD"] = np.sqrt(tidy.predVar)
        tidy["predUsefulness"] = tidy.predMean + stdSelInt * tidy.predSD
    elif predTheVars:
        tidy = pv[pv.Trait1 == pv.Trait2].rename(columns={"Trait1": "Trait"})[["sireID", "damID", "Nsegsnps", "predOf", "Trait", "predVar"]]
        tidy = tidy.assign(predSD=np.sqrt(tidy.predVar))
    else:
        tidy = pm.rename(columns={"Trait1": "Trait"})[["sireID", "damID", "predOf", "Trait", "predMean"]]
    return {"tidyPreds": tidy.reset_index(drop=True), "rawPreds": raw}


def predictCrossVars(modelType, snpeffs, parentfolds, haploMat, recombFreqMat, ncores=1, nBLASthreads=None):
    """R/parentWiseCrossVal.R:386-493: the cross-validation entry.  `snpeffs`: DataFrame with columns
    Repeat, Fold and `effects` (a DataFrame with Dataset, Trait and *snpeff list-columns);
    `parentfolds`: DataFrame with Repeat, Fold, CrossesToPredict.  All (Repeat, Fold) sets run against
    ONE device-resident session (haploMat / recombFreqMat are uploaded once), and the sets that are to be
    predicted for the SAME cross list go through one gms_predict_batched call (cross-side device state staged once)."""
    if modelType not in ("A", "AD", "DirDom"):
        raise ValueError("modelType must be 'A', 'AD' or 'DirDom'")
    pf = parentfolds.set_index(["Repeat", "Fold"])
    allparents = pd.unique(pd.concat([pd.concat([c["sireID"], c["damID"]]) for c in parentfolds["CrossesToPredict"]]).astype(str))
    session = CrossVarSession(haploMat, recombFreqMat, parents=allparents)
    comps = {"A": ["VarBV"], "AD": ["VarBV", "VarDD"], "DirDom": ["VarA", "VarD", "VarBV"]}[modelType]
    cols = {"A": ("allelesubsnpeff", None, None), "AD": ("allelesubsnpeff", "domdevsnpeff", None),
            "DirDom": ("addsnpeff", "domsnpeff", "allelesubsnpeff")}[modelType]
    jobs = []
    for pos, (_, row) in enumerate(snpeffs.iterrows()):
        eff = row["effects"]
        eff = eff[eff["Dataset"] == "trainset"]
        crosses = pf.loc[(row["Repeat"], row["Fold"]), "CrossesToPredict"]
        mats, traits = [], None
        for c in cols:
            if c is None:
                mats.append(None)
                continue
            t, mat = _effect_matrix(_effect_lists(eff, c), session.snp_names)
            if traits is not None and t != traits:
                raise ValueError("effect classes must name the same traits in the same order")
            traits = t
            mats.append(mat)
        sire, dam = session.indices(crosses["sireID"]), session.indices(crosses["damID"])
        jobs.append(dict(pos=pos, row=row, crosses=crosses, traits=traits, mats=mats, sire=sire, dam=dam,
                         key=(tuple(traits), sire.tobytes(), dam.tobytes())))
    out = [None] * len(jobs)
    try:
        groups = {}
        for j in jobs:
            groups.setdefault(j["key"], []).append(j)
        for grp in groups.values():
            start = time.perf_counter()
            stack = lambda i: None if grp[0]["mats"][i] is None else np.stack([j["mats"][i] for j in grp])
            o, nseg = session.engine.predict_batched(modelType, stack(0), stack(1), stack(2), grp[0]["sire"], grp[0]["dam"])
            elapsed = (time.perf_counter() - start) / len(grp)
            for si, j in enumerate(grp):
                tid = _tidy(j["crosses"], j["traits"], o[si], nseg, comps, elapsed)
                if modelType == "DirDom":      # VarBV rows first (R/parentWiseCrossVal.R:484-488)
                    tid["predVars"] = [pd.concat([p[p.predOf == "VarBV"], p[p.predOf != "VarBV"]], ignore_index=True)
                                       for p in tid["predVars"]]
                print(f"Done predicting fam vars. Took {round(elapsed / 60, 2)} mins for {len(tid)} crosses")
                out[j["pos"]] = {"Repeat": j["row"]["Repeat"], "Fold": j["row"]["Fold"],
                                 "modelType": j["row"].get("modelType", modelType), "predVars": tid}
    finally:
        session.close()
    return pd.DataFrame(out)
