"""GPU tests of the pieces around the contraction that round 2 added: the per-handle cache, the parent-sharded
two-phase compute with its exchange buffer, cross-validation batching (gms_predict_batched) and the device-side tidy
table (gms_tidy). Every number is checked against the oracle."""
import ctypes as C

import numpy as np
import pytest

from oracle import predcrossvar_oracle as O
from tests.util import oracle_slab, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


def synth(seed, m_c, P, T):
    rng = np.random.default_rng(seed)
    m_c = np.asarray(m_c)
    m = int(m_c.sum())
    cM = np.concatenate([np.sort(rng.uniform(0, 120, k)) for k in m_c])
    chrom = np.repeat(np.arange(1, len(m_c) + 1), m_c)
    p = rng.uniform(0.05, 0.95, m)
    H = (rng.random((2 * P, m)) < p).astype(np.uint8)
    add = rng.standard_normal((m, T))
    dom = 0.5 * rng.standard_normal((m, T))
    asub = add + dom * (1 - 2 * p)[:, None]
    R = O.recombfreq_to_ld_decay(O.genmap2recombfreq(cM, chrom))
    return dict(m=m, m_c=m_c, cM=cM, chrom=chrom, H=H, add=add, dom=dom, asub=asub, R=R, p=p)


@pytest.mark.parametrize("mode", ["auto", "fp64", "map_scan"])
def test_cache_reuses_tables_and_only_computes_new_parents(mode):
    import genomicmateselectr_b200 as g
    s = synth(61, [210, 150, 90], P=20, T=3)
    rng = np.random.default_rng(0)
    sire1, dam1 = rng.integers(0, 10, 60).astype(np.int32), rng.integers(0, 10, 60).astype(np.int32)     # parents 0-9
    sire2, dam2 = rng.integers(5, 20, 80).astype(np.int32), rng.integers(5, 20, 80).astype(np.int32)     # parents 5-19
    with g.CrossVarEngine(0) as eng:
        eng.set_genmap(s["cM"], s["m_c"])
        eng.set_haplotypes(s["H"])
        eng.set_mode(mode)
        a, na = eng.predict("AD", s["add"], s["dom"], None, sire1, dam1)
        st = eng.stats()
        assert st["last_cache_hit"] == 0 and st["last_parents_computed"] == len(set(sire1) | set(dam1))
        b, nb = eng.predict("AD", s["add"], s["dom"], None, sire1, dam1)          # same effects, same crosses
        st = eng.stats()
        assert st["last_cache_hit"] == 1 and st["last_parents_computed"] == 0
        assert np.array_equal(a, b) and np.array_equal(na, nb)
        c, nc = eng.predict("AD", s["add"], s["dom"], None, sire2, dam2)          # same effects, other crosses
        st = eng.stats()
        new = (set(sire2) | set(dam2)) - (set(sire1) | set(dam1))
        assert st["last_cache_hit"] == 1 and st["last_parents_computed"] == len(new)
        want, wseg = oracle_slab("AD", s["H"], s["R"], s["add"], s["dom"], None, sire2, dam2)
        assert np.array_equal(nc, wseg) and rel_err(c, want) < TOL
        dom2 = s["dom"].copy()
        dom2[7, 1] += 1e-3                                                        # one effect changes: everything is rebuilt
        d, nd = eng.predict("AD", s["add"], dom2, None, sire2, dam2)
        st = eng.stats()
        assert st["last_cache_hit"] == 0 and st["last_parents_computed"] == len(set(sire2) | set(dam2))
        want, _ = oracle_slab("AD", s["H"], s["R"], s["add"], dom2, None, sire2, dam2)
        assert rel_err(d, want) < TOL
        e, _ = eng.predict("A", s["add"], None, None, sire2, dam2)                # other model
        assert eng.stats()["last_cache_hit"] == 0
        want, _ = oracle_slab("A", s["H"], s["R"], s["add"], None, None, sire2, dam2)
        assert rel_err(e, want) < TOL
        eng.set_cache(False)
        f, _ = eng.predict("A", s["add"], None, None, sire2, dam2)
        assert eng.stats()["last_cache_hit"] == 0 and np.array_equal(e, f)
        # means in between must not disturb the cached operand images
        eng.set_cache(True)
        h1, _ = eng.predict("AD", s["add"], s["dom"], None, sire1, dam1)
        eng.cross_means(s["asub"], s["dom"], sire1, dam1)
        h2, _ = eng.predict("AD", s["add"], s["dom"], None, sire1, dam1)
        assert eng.stats()["last_cache_hit"] == 1 and np.array_equal(h1, h2) and np.array_equal(h1, a)


@pytest.mark.parametrize("mode", ["auto", "fp64"])
@pytest.mark.parametrize("model", ["AD", "DirDom"])
def test_parent_sharded_two_phase_compute(mode, model):
    """The multi-process multi-GPU protocol of the C ABI on ONE GPU: two handles stand for two ranks, each stages the whole
    cross list with its own range, evaluates its slice of the per-parent tables, the slices are exchanged through the
    buffers gms_parent_exchange exposes (an all-gather, done here with two device copies), then each runs its crosses."""
    import torch
    import genomicmateselectr_b200 as g
    s = synth(62, [300, 170, 260, 90], P=21, T=3)
    rng = np.random.default_rng(1)
    N = 333
    sire, dam = rng.integers(0, 21, N).astype(np.int32), rng.integers(0, 21, N).astype(np.int32)
    asub = s["asub"] if model == "DirDom" else None
    engines = [g.CrossVarEngine(0) for _ in range(2)]
    try:
        cut = 150
        ranges = [(0, cut), (cut, N)]
        for r, e in enumerate(engines):
            e.set_genmap(s["cM"], s["m_c"])
            e.set_haplotypes(s["H"])
            e.set_mode(mode)
            e.set_parent_shard(r, 2)
            e.stage_range(model, s["add"], s["dom"], asub, sire, dam, *ranges[r])
            e.compute_parents()
            with pytest.raises(g.GmsError, match="parent-sharded"):
                e.compute()
        torch.cuda.synchronize()
        bufs = [e.parent_exchange() for e in engines]
        assert bufs[0][1] == bufs[1][1] > 0 and bufs[0][2] == 2
        n = bufs[0][1] * 8
        cudart = C.CDLL("libcudart.so")
        for r in range(2):             # rank r receives the other rank's slice
            o = 1 - r
            rc = cudart.cudaMemcpy(C.c_void_p(bufs[r][0] + o * n), C.c_void_p(bufs[o][0] + o * n), C.c_size_t(n), 3)
            assert rc == 0
        outs = []
        for r, e in enumerate(engines):
            e.compute_crosses()
            assert e.stats()["last_parents_computed"] in (10, 11)
            outs.append(e.fetch())
        out = np.concatenate([o for o, _ in outs])
        nseg = np.concatenate([n_ for _, n_ in outs])
    finally:
        for e in engines:
            e.close()
    want, wseg = oracle_slab(model, s["H"], s["R"], s["add"], s["dom"], asub, sire, dam)
    assert np.array_equal(nseg, wseg) and rel_err(out, want) < TOL
    with g.CrossVarEngine(0) as one:
        one.set_genmap(s["cM"], s["m_c"])
        one.set_haplotypes(s["H"])
        one.set_mode(mode)
        ref, _ = one.predict(model, s["add"], s["dom"], asub, sire, dam)
    assert np.array_equal(out, ref)            # sharding changes who computes a table row, not its value


@pytest.mark.parametrize("model", ["A", "AD", "DirDom"])
def test_predict_batched_cv_folds(model):
    """gms_predict_batched: 4 effect sets (cross-validation folds, R/parentWiseCrossVal.R:430-490) against one cross list:
    every set against the oracle, and bit-identical to 4 separate gms_predict calls."""
    import genomicmateselectr_b200 as g
    s = synth(63, [190, 140, 66], P=12, T=2)
    rng = np.random.default_rng(2)
    nsets, N = 4, 150
    sire, dam = rng.integers(0, 12, N).astype(np.int32), rng.integers(0, 12, N).astype(np.int32)
    adds = np.stack([s["add"] + 0.3 * rng.standard_normal(s["add"].shape) for _ in range(nsets)])
    doms = np.stack([s["dom"] + 0.1 * rng.standard_normal(s["dom"].shape) for _ in range(nsets)])
    asubs = adds + 0.2 * doms
    d = doms if model != "A" else None
    a = asubs if model == "DirDom" else None
    with g.CrossVarEngine(0) as eng:
        eng.set_genmap(s["cM"], s["m_c"])
        eng.set_haplotypes(s["H"])
        out, nseg = eng.predict_batched(model, adds, d, a, sire, dam)
        assert out.shape[:2] == (nsets, N)
        for k in range(nsets):
            want, wseg = oracle_slab(model, s["H"], s["R"], adds[k], None if d is None else d[k], None if a is None else a[k], sire, dam)
            assert np.array_equal(nseg, wseg) and rel_err(out[k], want) < TOL
            single, _ = eng.predict(model, adds[k], None if d is None else d[k], None if a is None else a[k], sire, dam)
            assert np.array_equal(single, out[k])


@pytest.mark.parametrize("model", ["A", "AD", "DirDom"])
@pytest.mark.parametrize("selind", [False, True])
def test_device_tidy_table(model, selind):
    """gms_tidy (TGV, SELIND, predSD, usefulness fused on the device) against the same quantities built from the oracle's
    variances and means the way R/gmsFunctions.R:726-871 builds them."""
    import genomicmateselectr_b200 as g
    T = 3
    s = synth(64, [160, 120], P=9, T=T)
    rng = np.random.default_rng(3)
    N = 40
    sire, dam = rng.integers(0, 9, N).astype(np.int32), rng.integers(0, 9, N).astype(np.int32)
    w = np.array([0.5, 0.3, 0.2]) if selind else None
    i_sel = 2.67
    # predictCrosses' effect classes: A / AD: AddEffectList = allele substitution effects; DirDom: (add, dom) + asub
    add, dom, asub = s["add"], (s["dom"] if model != "A" else None), (s["asub"] if model == "DirDom" else None)
    with g.CrossVarEngine(0) as eng:
        eng.set_genmap(s["cM"], s["m_c"])
        eng.set_haplotypes(s["H"])
        out, nseg = eng.predict(model, add, dom, asub, sire, dam)
        tab = eng.tidy(w, i_sel)
    want, _ = oracle_slab(model, s["H"], s["R"], add, dom, asub, sire, dam)
    pairs = [(t, t) for t in range(T)] + [(a, b) for a in range(T) for b in range(a + 1, T)]
    dose = s["H"][0::2].astype(float) + s["H"][1::2]
    bv_eff = asub if model == "DirDom" else add
    traits = [f"t{i}" for i in range(T)]
    nof = 1 if model == "A" else 2
    assert tab.shape == (N, nof, T + int(selind), 4)
    for x in range(N):
        gebv = dose @ bv_eff
        mean_bv = (gebv[sire[x]] + gebv[dam[x]]) / 2
        if model == "DirDom":
            mean_tgv = np.array([O.predCrossMeans_TGV(dose[sire[x]], dose[dam[x]], add[:, t], dom[:, t]) for t in range(T)])
        else:
            mean_tgv = mean_bv
        for of in range(nof):
            if of == 0:
                var = want[x, :, 2 if model == "DirDom" else 0]
                mean = mean_bv
            else:
                var = want[x, :, 0] + want[x, :, 1]
                mean = mean_tgv
            rows = []
            if selind:
                G = np.zeros((T, T))
                for i, (a, b) in enumerate(pairs):
                    G[a, b] = G[b, a] = var[i]
                rows.append((w @ mean, w @ G @ w))
            rows += [(mean[t], var[t]) for t in range(T)]
            for slot, (mu, v) in enumerate(rows):
                got = tab[x, of, slot]
                assert got[0] == pytest.approx(mu, rel=1e-10, abs=1e-10)
                assert got[1] == pytest.approx(v, rel=1e-9)
                assert got[2] == pytest.approx(np.sqrt(v), rel=1e-9)
                assert got[3] == pytest.approx(mu + i_sel * np.sqrt(v), rel=1e-9, abs=1e-9)
