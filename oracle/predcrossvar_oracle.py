"""CPU ORACLE (test infrastructure, not product code).

A NumPy float64 restatement of genomicMateSelectR's cross-variance hot path in
the REFERENCE'S OWN FORMULATION: per cross it subsets to the segregating SNPs,
materialises both parents' gametic LD matrices, sums them to the m_seg x m_seg
cross LD matrix `D`, and evaluates one quadratic form per trait pair (plus a
fresh `D*D` per trait pair for dominance).  It deliberately does NOT use the
delta / heterozygosity reformulation the CUDA kernels use, so that it is an
independent check of that algebra.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.  The product package
(`genomicmateselectr_b200`) never does; it fails loudly without its CUDA library.

Parity pins: the reference has no test-suite; its only known answers are the
printed outputs of its two vignettes (3-4 significant figures).  This oracle is
pinned to every one of those for modelType A and AD in
`tests/test_oracle_golden.py` (bundled data, parents 49/153/74/228/146).
DirDom is "parity unpinned by the reference" (the effects behind the printed
DirDom numbers are not stored in the package) - for DirDom
the oracle is only a restatement of the code.

R is not installed here, so this follows the R source line by line instead of
calling it.  Citations are `file:line` under /root/reference.
"""
from __future__ import annotations

import itertools
import time

import numpy as np

__all__ = [
    "genmap2recombfreq", "recombfreq_to_ld_decay", "crosses2predict", "quadform",
    "calcGameticLD", "calcCrossLD", "predOneCrossVar", "predOneCross", "predCrossVars",
    "trait_pairs", "predCrossMeans_BV", "predCrossMeans_TGV", "selind_var",
]


# --------------------------------------------------------------------------- helpers.R
def genmap2recombfreq(m, chrom):
    """R/helpers.R:204-213.  `m`: cM positions; `chrom`: integer chromosome of each
    SNP (the reference derives it by regex on the "<chr>_<id>" names, :208-211).
    Returns c1 (pairwise Haldane recombination frequency; 0.5 across chromosomes)."""
    m = np.asarray(m, dtype=np.float64)
    chrom = np.asarray(chrom)
    d = np.abs(m[:, None] - m[None, :])               # dist(m, method="manhattan")  :205
    c1 = 0.5 * (1.0 - np.exp(-2.0 * (d / 100.0)))     # :206
    off = chrom[:, None] != chrom[None, :]            # :208-211
    c1[off] = 0.5
    return c1


def recombfreq_to_ld_decay(c1):
    """The caller-side `recombFreqMat <- 1-(2*recombFreqMat)`
    (vignettes/start_here.Rmd:112; data-raw/simulate_example_data.R:111)."""
    return 1.0 - (2.0 * c1)


def crosses2predict(parents):
    """R/helpers.R:278-289: upper triangle incl. diagonal of the parents x parents
    mating matrix, row-major over (sire, dam) after pivot_longer."""
    parents = list(parents)
    return [(parents[i], parents[j]) for i in range(len(parents)) for j in range(i, len(parents))]


def quadform(D, x, y, longdouble=True):
    """R/helpers.R:302 `as.numeric(colSums(x*(D%*%y)))`.  R's colSums accumulates
    in long double; `longdouble=False` gives the plain float64 pairwise sum."""
    v = x * (D @ y)
    if longdouble:
        return float(np.sum(v.astype(np.longdouble)))
    return float(np.sum(v))


# --------------------------------------------------------------------------- predCrossVar.R
def calcGameticLD(X, R):
    """R/predCrossVar.R:261-265.  X: the parent's 2 x m haplotype rows (HapA, HapB)."""
    X = np.asarray(X, dtype=np.float64)
    p = X.mean(axis=0)                                 # colMeans(X)           :263
    return R * ((0.5 * X.T) @ X - np.outer(p, p))      # :264


def calcCrossLD(Xsire, Xdam, R):
    """R/predCrossVar.R:280-282."""
    return calcGameticLD(Xsire, R) + calcGameticLD(Xdam, R)


def trait_pairs(traits):
    """Row order of a cross's result: T variances (Ti,Ti) first, then
    combn(traits,2) in column order (1,2),(1,3),...,(T-1,T).  R/predCrossVar.R:128-138."""
    traits = list(traits)
    pairs = [(t, t) for t in traits]
    if len(traits) > 1:
        pairs += list(itertools.combinations(traits, 2))
    return pairs


def predOneCrossVar(t1, t2, D, modelType, add, dom=None, longdouble=True):
    """R/predCrossVar.R:180-247, VPM branch.  Returns [("VarA", v)] or
    [("VarA", v), ("VarD", v)].  `D*D` is formed per trait pair, as the reference does (:219)."""
    out = [("VarA", quadform(D, add[t1], add[t2], longdouble))]            # :209-211
    if modelType == "AD":
        D2 = D * D                                                           # :219
        out.append(("VarD", quadform(D2, dom[t1], dom[t2], longdouble)))   # :220-222
    return out


def predOneCross(sire_rows, dam_rows, modelType, H, R, add, dom=None, longdouble=True):
    """R/predCrossVar.R:96-162 on arrays.  `sire_rows`/`dam_rows`: the two row indices
    (HapA, HapB) of each parent in H (n_rows x m, values {0,1}); R: m x m LD-decay
    matrix; add/dom: dict trait -> length-m effect vector.
    Returns (Nsegsnps, [(Trait1, Trait2, predOf, predVar), ...])."""
    Xs = H[list(sire_rows), :].astype(np.float64)
    Xd = H[list(dam_rows), :].astype(np.float64)
    x = Xs.sum(axis=0) + Xd.sum(axis=0)                # colSums(rbind(...))     :111-112
    seg = np.flatnonzero((x > 0) & (x < 4))            # :113
    if seg.size == 0:                                  # :156-160 (cross vanishes on unnest)
        return 0, []
    Rs = R[np.ix_(seg, seg)]                           # :115
    adds = {t: np.asarray(v, dtype=np.float64)[seg] for t, v in add.items()}       # :117
    doms = None
    if modelType == "AD":
        doms = {t: np.asarray(v, dtype=np.float64)[seg] for t, v in dom.items()}   # :118-119
    D = calcCrossLD(Xs[:, seg], Xd[:, seg], Rs)        # :123
    rows = []
    for t1, t2 in trait_pairs(list(add.keys())):       # :128-151
        for pred_of, v in predOneCrossVar(t1, t2, D, modelType, adds, doms, longdouble):
            rows.append((t1, t2, pred_of, v))
    return int(seg.size), rows


def predCrossVars(crosses, modelType, AddEffectList, DomEffectList=None, predType="VPM",
                  haploMat=None, hap_rownames=None, recombFreqMat=None, longdouble=True):
    """R/predCrossVar.R:24-78 on labelled arrays.

    crosses        : iterable of (sireID, damID) strings
    AddEffectList  : dict trait -> length-m vector (the reference passes 1 x m matrices
                     and recovers them through scale()/"scaled:center", :32-43 - for VPM
                     that is the identity)
    haploMat       : (2*n_ind) x m array in {0,1}; hap_rownames: "<GID>_HapA"/"<GID>_HapB"
    recombFreqMat  : m x m, already 1-2c
    Returns a list of dicts (one per cross that has >= 1 segregating SNP, in input
    order): sireID, damID, Nsegsnps, ComputeTime, predVars=[(Trait1,Trait2,predOf,predVar)].
    """
    if predType != "VPM":
        raise ValueError("only predType='VPM' is restated (PMV is broken in the reference; SURVEY 8a/a6)")
    H = np.asarray(haploMat)
    R = np.asarray(recombFreqMat, dtype=np.float64)
    row_of = {n: i for i, n in enumerate(hap_rownames)}
    eff_model = "AD" if modelType == "AD" else "A"     # anything else behaves as "A"  :33,39,119,216
    out = []
    for sire, dam in crosses:
        t0 = time.perf_counter()
        try:
            s_rows = (row_of[f"{sire}_HapA"], row_of[f"{sire}_HapB"])
            d_rows = (row_of[f"{dam}_HapA"], row_of[f"{dam}_HapB"])
        except KeyError as e:                          # R: subscript out of bounds at :46
            raise KeyError(f"parent haplotype row {e} not in haploMat") from None
        nseg, rows = predOneCross(s_rows, d_rows, eff_model, H, R, AddEffectList,
                                  DomEffectList if eff_model == "AD" else None, longdouble)
        if nseg == 0:
            continue
        out.append(dict(sireID=sire, damID=dam, Nsegsnps=nseg,
                        ComputeTime=time.perf_counter() - t0, predVars=rows))
    return out


# --------------------------------------------------------------------------- "next" rows (8f)
def predCrossMeans_BV(dose_sire, dose_dam, add_t):
    """R/predCrossVar.R:354-370: mid-parent GEBV, GEBV = doseMat %*% effect."""
    s = float(np.dot(np.asarray(dose_sire, dtype=np.float64), add_t))
    d = float(np.dot(np.asarray(dose_dam, dtype=np.float64), add_t))
    return s, d, (s + d) / 2.0


def predCrossMeans_TGV(dose_sire, dose_dam, add_t, dom_t):
    """R/predCrossVar.R:383-392 (Falconer & MacKay eqn 14.6)."""
    p1 = np.asarray(dose_sire, dtype=np.float64) / 2.0
    p2 = np.asarray(dose_dam, dtype=np.float64) / 2.0
    q = 1.0 - p1
    y = p1 - p2
    g = add_t * (p1 - q - y) + dom_t * ((2.0 * p1 * q) + y * (p1 - q))
    return float(np.sum(g))


def selind_var(pred_rows, traits, SIwts):
    """R/gmsFunctions.R:825-853: G from the (Trait1,Trait2,predVar) rows of one cross and
    one predOf, symmetrised from its upper triangle (:841), then SIwts' G SIwts.
    `SIwts`: dict trait -> weight (its key order picks rows/cols, :842)."""
    idx = {t: i for i, t in enumerate(traits)}
    G = np.full((len(traits), len(traits)), np.nan)
    for t1, t2, v in pred_rows:
        G[idx[t1], idx[t2]] = v
    low = np.tril_indices(len(traits), -1)
    G[low] = G.T[low]
    order = [idx[t] for t in SIwts]
    w = np.array([SIwts[t] for t in SIwts], dtype=np.float64)
    G = G[np.ix_(order, order)]
    return float(w @ G @ w)
