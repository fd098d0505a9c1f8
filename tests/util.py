"""Shared helpers for the parity tests: run the CPU oracle and shape its answer like the C ABI's slab."""
import numpy as np

from oracle import predcrossvar_oracle as O


def pair_index(T):
    pairs = [(t, t) for t in range(T)] + [(a, b) for a in range(T) for b in range(a + 1, T)]
    return pairs


def oracle_slab(model, H, R, add, dom, asub, sire, dam, longdouble=True):
    """out[N, npairs, ncomp], nseg[N] from the oracle (reference formulation, dense R)."""
    T = add.shape[1]
    traits = [f"t{i}" for i in range(T)]
    A = {t: add[:, i] for i, t in enumerate(traits)}
    D = {t: dom[:, i] for i, t in enumerate(traits)} if dom is not None else None
    S = {t: asub[:, i] for i, t in enumerate(traits)} if asub is not None else None
    ncomp = {"A": 1, "AD": 2, "DirDom": 3}[model]
    npairs = T * (T + 1) // 2
    out = np.zeros((len(sire), npairs, ncomp))
    nseg = np.zeros(len(sire), np.int32)
    for x, (s, d) in enumerate(zip(sire, dam)):
        rs, rd = (2 * s, 2 * s + 1), (2 * d, 2 * d + 1)
        n, rows = O.predOneCross(rs, rd, "AD" if model != "A" else "A", H, R, A, D, longdouble)
        nseg[x] = n
        if n == 0:
            continue
        vals = [v for _, _, _, v in rows]
        if model == "A":
            out[x, :, 0] = vals
        else:
            out[x, :, 0] = vals[0::2]
            out[x, :, 1] = vals[1::2]
        if model == "DirDom":     # second predCrossVars call of R/gmsFunctions.R:712-718
            _, rows2 = O.predOneCross(rs, rd, "A", H, R, S, None, longdouble)
            out[x, :, 2] = [v for _, _, _, v in rows2]
    return out, nseg


def oracle_slab_blockwise(model, H, blocks, offs, add, dom, asub, sire, dam):
    """Same answer computed chromosome by chromosome with the SAME oracle code and added up
    (recombFreqMat is exactly 0 across chromosomes, so D is block diagonal and every quadratic
    form is a sum over chromosomes). Lets a cassava-scale cross be checked in seconds."""
    tot = None
    nseg = None
    for R, (a, b) in zip(blocks, offs):
        o, n = oracle_slab(model, H[:, a:b], R, add[a:b], None if dom is None else dom[a:b],
                           None if asub is None else asub[a:b], sire, dam)
        tot = o if tot is None else tot + o
        nseg = n if nseg is None else nseg + n
    return tot, nseg


def oracle_slab_c(model, H, add, dom, asub, sire, dam, R=None, blocks=None, cM=None, chrom=None, nthreads=0):
    """out[N, npairs, ncomp], nseg[N] from the oracle's C restatement (oracle/crossvar_ref.c: the reference's dense
    m_seg x m_seg formulation, OpenMP) - for problem sizes where the NumPy oracle would take minutes per cross."""
    from oracle import crossvar_ref as CR
    ad = model != "A"
    prob = CR.RefProblem(H, add, dom if ad else None, R=R, blocks=blocks, cM=cM, chrom=chrom)
    prob2 = CR.RefProblem(H, asub, None, R=R, blocks=blocks, cM=cM, chrom=chrom) if model == "DirDom" else None
    T = add.shape[1]
    ncomp = {"A": 1, "AD": 2, "DirDom": 3}[model]
    out = np.zeros((len(sire), T * (T + 1) // 2, ncomp))
    nseg = np.zeros(len(sire), np.int32)
    for x, (s, d) in enumerate(zip(sire, dam)):
        o, n = prob.one_cross(int(s), int(d), "AD" if ad else "A", nthreads)
        out[x, :, : o.shape[1]] = o
        nseg[x] = n
        if prob2 is not None:      # second predCrossVars call of R/gmsFunctions.R:712-718
            o2, _ = prob2.one_cross(int(s), int(d), "A", nthreads)
            out[x, :, 2] = o2[:, 0]
    return out, nseg


def strict_rel_err(got, want):
    """max |got-want| / |want| over the VARIANCES (the first T pairs of every cross): no guard at all."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    T = int((np.sqrt(8 * want.shape[1] + 1) - 1) / 2)
    g, w = got[:, :T, :], want[:, :T, :]
    denom = np.where(w == 0, 1.0, np.abs(w))
    return float(np.max(np.abs(g - w) / denom))


def rel_err(got, want):
    """max |got-want| / |want| with the guard the north-star tolerance implies: entries that are
    tiny by cancellation are measured against sqrt(var_t1 * var_t2) of the same cross/component."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    T = int((np.sqrt(8 * want.shape[1] + 1) - 1) / 2)
    pairs = pair_index(T)
    var = np.abs(want[:, :T, :])
    scale = np.empty_like(want)
    for i, (a, b) in enumerate(pairs):
        scale[:, i, :] = np.sqrt(var[:, a, :] * var[:, b, :])
    denom = np.maximum(np.abs(want), 1e-6 * scale)
    denom = np.where(denom == 0, 1.0, denom)
    return float(np.max(np.abs(got - want) / denom))
