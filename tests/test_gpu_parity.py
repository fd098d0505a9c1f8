"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle.
Tolerance (BASELINE.json north_star): 1e-9 relative on every variance and covariance;
Nsegsnps bit-exact."""
import numpy as np
import pytest

from oracle import predcrossvar_oracle as O
from tests.util import oracle_slab, oracle_slab_blockwise, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def gms():
    import genomicmateselectr_b200 as g
    return g


@pytest.fixture(params=["auto", "fp64"])
def mode(request):
    """"auto" = the default (exact int8 tcgen05 path when T <= 6), "fp64" = the FP64 DMMA contraction."""
    return request.param


def bundled_engine(gms, b, mode="auto"):
    eng = gms.CrossVarEngine(0)
    eng.set_mode(mode)
    idx = []
    for p in b.PARENTS:
        idx += [b.hap_rownames.index(f"{p}_HapA"), b.hap_rownames.index(f"{p}_HapB")]
    H = b.haploMat[idx]
    eng.set_recomb_matrix(b.recombFreqMat, b.chrom)
    eng.set_haplotypes(H)
    sire, dam = zip(*[(i, j) for i in range(5) for j in range(i, 5)])
    return eng, H, np.array(sire, np.int32), np.array(dam, np.int32)


# ---------------------------------------------------------------- configs 1-3: bundled data
def test_bundled_A_single_trait(gms, bundled, mode):
    b = bundled
    eng, H, sire, dam = bundled_engine(gms, b, mode)
    out, nseg = eng.predict("A", b.effA[:, :1], None, None, sire, dam)
    want, wseg = oracle_slab("A", H, b.recombFreqMat, b.effA[:, :1], None, None, sire, dam)
    assert out.shape == (15, 1, 1)
    assert np.array_equal(nseg, wseg) and list(nseg[:4]) == [22, 46, 48, 53]
    assert rel_err(out, want) < TOL
    assert round(out[0, 0, 0]) == 707          # docs/articles/start_here.html:385
    eng.close()


def test_bundled_A_two_traits_matches_vignette(gms, bundled, mode):
    b = bundled
    eng, H, sire, dam = bundled_engine(gms, b, mode)
    out, nseg = eng.predict("A", b.effA, None, None, sire, dam)
    want, wseg = oracle_slab("A", H, b.recombFreqMat, b.effA, None, None, sire, dam)
    assert np.array_equal(nseg, wseg)
    assert rel_err(out, want) < TOL
    # slab pair order (T1,T1),(T2,T2),(T1,T2): 707 1519 -350 | 1079 1813 (start_here.html:385)
    assert [round(v) for v in out[0, :, 0]] == [707, 1519, -350]
    assert [round(v) for v in out[1, :2, 0]] == [1079, 1813]
    si = eng.selind("A", out, [0.75, 0.25])
    assert [round(v) for v in si[:4, 0]] == [361, 454, 458, 536]      # start_here.html:333-342
    eng.close()


def test_bundled_AD_matches_vignette(gms, bundled, mode):
    b = bundled
    eng, H, sire, dam = bundled_engine(gms, b, mode)
    out, nseg = eng.predict("AD", b.effAD_add, b.effAD_dom, None, sire, dam)
    want, wseg = oracle_slab("AD", H, b.recombFreqMat, b.effAD_add, b.effAD_dom, None, sire, dam)
    assert np.array_equal(nseg, wseg)
    assert rel_err(out, want) < TOL
    x = 14                                         # 146 x 146 is the last cross
    assert nseg[x] == 34
    assert [round(v) for v in out[x, :, 0]] == [1232, 1976, -820]     # non_additive_models.html:624-626
    assert [round(v, 2) for v in out[x, :, 1]] == [9.97, 51.27, 9.04]  # :627-629 (51.3 printed)
    si = eng.selind("AD", out, [0.75, 0.25])
    assert round(si[x, 0]) == 509 and round(si[x, 0] + si[x, 1]) == 521   # :600-601
    eng.close()


def test_bundled_DirDom(gms, bundled, mode):
    b = bundled
    a, d, alpha = b.dirdom_effects()
    eng, H, sire, dam = bundled_engine(gms, b, mode)
    out, nseg = eng.predict("DirDom", a, d, alpha, sire, dam)
    want, wseg = oracle_slab("DirDom", H, b.recombFreqMat, a, d, alpha, sire, dam)
    assert out.shape == (15, 3, 3)
    assert np.array_equal(nseg, wseg)
    assert rel_err(out, want) < TOL
    eng.close()


# ---------------------------------------------------------------- synthetic, ragged shapes
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
    return dict(m=m, m_c=m_c, cM=cM, chrom=chrom, H=H, add=add, dom=dom, asub=asub, R=R)


@pytest.mark.parametrize("model", ["A", "AD", "DirDom"])
@pytest.mark.parametrize("T", [1, 2, 3, 5, 7, 16])
def test_synthetic_map_path(gms, model, T, mode):
    s = synth(100 + T, [200, 131, 77, 300], P=24, T=T)
    rng = np.random.default_rng(5)
    sire = rng.integers(0, 24, 70).astype(np.int32)
    dam = rng.integers(0, 24, 70).astype(np.int32)
    sire[:6] = dam[:6]                      # selfs
    sire[6], dam[6] = sire[7], dam[7]       # a duplicate row
    eng = gms.CrossVarEngine(0)
    eng.set_mode(mode)
    eng.set_genmap(s["cM"], s["m_c"])
    eng.set_haplotypes(s["H"])
    dom = s["dom"] if model != "A" else None
    asub = s["asub"] if model == "DirDom" else None
    out, nseg = eng.predict(model, s["add"], dom, asub, sire, dam)
    want, wseg = oracle_slab(model, s["H"], s["R"], s["add"], dom, asub, sire, dam)
    assert np.array_equal(nseg, wseg)
    assert rel_err(out, want) < TOL
    assert np.array_equal(out[6], out[7])
    eng.close()


def test_blocks_path_equals_map_path_and_many_rowtiles(gms, mode):
    s = synth(11, [700, 390, 129], P=12, T=5)     # 6 + 4 + 2 row tiles; ragged last tiles
    sire = np.array([0, 1, 2, 3, 4, 5, 6, 7, 3], np.int32)
    dam = np.array([1, 1, 5, 9, 10, 11, 6, 0, 8], np.int32)
    eng = gms.CrossVarEngine(0)
    eng.set_mode(mode)
    eng.set_genmap(s["cM"], s["m_c"])
    eng.set_haplotypes(s["H"])
    out1, nseg1 = eng.predict("AD", s["add"], s["dom"], None, sire, dam)
    want, wseg = oracle_slab("AD", s["H"], s["R"], s["add"], s["dom"], None, sire, dam)
    assert np.array_equal(nseg1, wseg) and rel_err(out1, want) < TOL
    eng2 = gms.CrossVarEngine(0)
    eng2.set_mode(mode)
    assert eng2.set_recomb_matrix(s["R"], s["chrom"]) is True
    eng2.set_haplotypes(s["H"].astype(np.int32).copy(order="F"))   # R integer-matrix layout entry point
    out2, nseg2 = eng2.predict("AD", s["add"], s["dom"], None, sire, dam)
    assert np.array_equal(nseg2, wseg) and rel_err(out2, want) < TOL
    # chromosome shards add up to the whole (multi-GPU chromosome axis)
    acc = np.zeros_like(out1)
    accn = np.zeros_like(nseg1)
    for c in range(3):
        eng.set_chromosome_shard(c, c + 1)
        o, n = eng.predict("AD", s["add"], s["dom"], None, sire, dam)
        acc += o
        accn += n
    assert np.array_equal(accn, wseg) and rel_err(acc, want) < TOL
    eng.close()
    eng2.close()


def test_general_matrix_single_block(gms, mode):
    """recombFreqMat with non-zero off-chromosome entries -> one dense block."""
    rng = np.random.default_rng(3)
    m, P, T = 150, 6, 2
    B = rng.standard_normal((m, m))
    R = B @ B.T / m
    chrom = np.repeat([1, 2], [80, 70])
    H = (rng.random((2 * P, m)) < 0.5).astype(np.uint8)
    add, dom = rng.standard_normal((m, T)), rng.standard_normal((m, T))
    sire = np.array([0, 1, 2, 3], np.int32)
    dam = np.array([0, 2, 4, 5], np.int32)
    eng = gms.CrossVarEngine(0)
    eng.set_mode(mode)
    assert eng.set_recomb_matrix(R, chrom) is False
    eng.set_haplotypes(H)
    out, nseg = eng.predict("AD", add, dom, None, sire, dam)
    want, wseg = oracle_slab("AD", H, R, add, dom, None, sire, dam)
    assert np.array_equal(nseg, wseg) and rel_err(out, want) < TOL
    eng.close()


# ---------------------------------------------------------------- edge cases and errors
def test_edge_cases(gms, mode):
    s = synth(2, [64, 40], P=4, T=2)
    H = s["H"].copy()
    H[0] = H[1] = H[2] = H[3] = 1            # parents 0 and 1: fixed for allele 1 everywhere
    eng = gms.CrossVarEngine(0)
    eng.set_mode(mode)
    eng.set_genmap(s["cM"], s["m_c"])
    eng.set_haplotypes(H)
    out, nseg = eng.predict("AD", s["add"], s["dom"], None, [0, 0, 2], [1, 0, 3])
    assert list(nseg[:2]) == [0, 0] and np.all(out[:2] == 0.0)     # vanishing crosses (R/predCrossVar.R:156-160)
    want, wseg = oracle_slab("AD", H, s["R"], s["add"], s["dom"], None, [2], [3])
    assert nseg[2] == wseg[0] and rel_err(out[2:], want) < TOL
    out0, nseg0 = eng.predict("A", s["add"], None, None, [], [])   # empty cross list
    assert out0.shape == (0, 3, 1) and nseg0.shape == (0,)
    with pytest.raises(gms.GmsError, match="out of range"):
        eng.predict("A", s["add"], None, None, [0], [9])
    with pytest.raises(gms.GmsError, match="required"):
        eng.predict("AD", s["add"], None, None, [0], [1])
    nan_dom = s["dom"].copy()
    nan_dom[7, 0] = np.nan
    with pytest.raises(gms.GmsError, match="non-finite marker effect"):      # the reference would return NaN variances
        eng.predict("AD", s["add"], nan_dom, None, [0], [1])
    bad = H.copy()
    bad[3, 5] = 2
    with pytest.raises(gms.GmsError, match="outside"):
        eng.set_haplotypes(bad)
    eng.close()
    eng = gms.CrossVarEngine(0)
    with pytest.raises(gms.GmsError, match="before"):
        eng.set_haplotypes(H)
    R = s["R"].copy()
    R[3, 10] += 1e-3
    with pytest.raises(gms.GmsError, match="symmetric"):
        eng.set_recomb_matrix(R, s["chrom"])
    eng.close()


def test_invariances(gms, mode):
    """Size-independent properties of the path: sire/dam swap, HapA/HapB swap, effect scaling."""
    s = synth(9, [300, 200], P=10, T=3)
    sire = np.arange(10, dtype=np.int32)
    dam = np.roll(sire, 3)
    eng = gms.CrossVarEngine(0)
    eng.set_mode(mode)
    eng.set_genmap(s["cM"], s["m_c"])
    eng.set_haplotypes(s["H"])
    base, nseg = eng.predict("AD", s["add"], s["dom"], None, sire, dam)
    swapped, nseg2 = eng.predict("AD", s["add"], s["dom"], None, dam, sire)
    assert np.array_equal(nseg, nseg2) and np.allclose(base, swapped, rtol=1e-12, atol=0)
    scaled, _ = eng.predict("AD", 3.0 * s["add"], -2.0 * s["dom"], None, sire, dam)
    assert np.allclose(scaled[:, :, 0], 9.0 * base[:, :, 0], rtol=1e-12)
    assert np.allclose(scaled[:, :, 1], 4.0 * base[:, :, 1], rtol=1e-12)
    H2 = s["H"].copy()
    H2[[0, 1]] = H2[[1, 0]]                  # swap the two haplotypes of parent 0
    eng.set_haplotypes(H2)
    flipped, nseg3 = eng.predict("AD", s["add"], s["dom"], None, sire, dam)
    assert np.array_equal(nseg, nseg3) and np.allclose(base, flipped, rtol=1e-11, atol=0)
    eng.close()


# ---------------------------------------------------------------- BASELINE.json full size
@pytest.fixture(scope="module")
def cassava():
    """configs[3]: 1000 parents, 33k SNPs / 18 chromosomes, 5 traits, AD. 40 crosses of the benchmark's own 50,000
    (32 random ones, the first, the last, and the first selfs) evaluated by the oracle's C restatement of the reference
    algorithm (dense m_seg x m_seg cross LD matrix; matrix entries from the map, R/helpers.R:204-213)."""
    from genomicmateselectr_b200 import synth as S
    from tests.util import oracle_slab_c
    d = S.make("cassava")
    rng = np.random.default_rng(2026)
    selfs = np.flatnonzero(d["sire"] == d["dam"])[:6]
    pick = np.unique(np.concatenate([rng.choice(d["N"], 32, replace=False), [0, d["N"] - 1], selfs]))
    offs = np.concatenate([[0], np.cumsum(d["m_c"])])
    blocks = [O.recombfreq_to_ld_decay(O.genmap2recombfreq(d["cM"][a:b], np.ones(b - a))) for a, b in zip(offs[:-1], offs[1:])]
    want, wseg = oracle_slab_c("AD", d["H"], d["add"], d["dom"], None, d["sire"][pick], d["dam"][pick], blocks=blocks)
    return d, pick, want, wseg


def test_cassava_scale_against_the_oracle(gms, cassava, mode):
    """The whole 50,000-cross batch of the benchmark in both modes; the sampled crosses against the oracle, the whole
    batch through invariants (finite, positive variances, duplicates of a sampled cross identical)."""
    d, pick, want, wseg = cassava
    assert len(pick) >= 34
    eng = gms.CrossVarEngine(0)
    eng.set_mode(mode)
    eng.set_genmap(d["cM"], d["m_c"])
    eng.set_haplotypes(d["H"])
    N = d["N"] if mode == "auto" else 4000                  # the FP64 contraction is 12x slower: a 4,000-cross batch
    idx = np.concatenate([pick, np.setdiff1d(np.arange(N), pick)[: N - len(pick)]])
    out, nseg = eng.predict("AD", d["add"], d["dom"], None, d["sire"][idx], d["dam"][idx])
    st = eng.stats()
    assert np.all(np.isfinite(out)) and np.all(nseg > 0)
    assert np.all(out[:, :5, :] > 0)                       # variances are positive
    assert np.array_equal(nseg[: len(pick)], wseg)
    assert rel_err(out[: len(pick)], want) < TOL
    from tests.util import strict_rel_err
    assert strict_rel_err(out[: len(pick)], want) < TOL
    if mode == "auto":
        assert st["last_mode"] == 2 and st["last_fallback_units"] == 0     # N(0,1) effects: nothing fails the error bound
    eng.close()


@pytest.fixture(scope="module")
def large_geometry():
    """configs[4] geometry: chromosomes of 5,556 and 5,555 SNPs (87 k-blocks per row), T = 10 (the trait-outer kernel
    template), parent ids spread over a 5,000-parent index space; 10 crosses against the oracle's C restatement."""
    from tests.util import oracle_slab_c
    rng = np.random.default_rng(20261119)
    m_c = np.array([5556, 5555])
    m, P, T = int(m_c.sum()), 5000, 10
    cM = np.concatenate([np.sort(rng.uniform(0.0, 130.0, int(k))) for k in m_c])
    chrom = np.repeat([1, 2], m_c).astype(np.int32)
    p = rng.uniform(0.05, 0.95, m).astype(np.float32)
    H = np.empty((2 * P, m), np.uint8)
    for a in range(0, 2 * P, 500):
        H[a:a + 500] = rng.random((500, m), dtype=np.float32) < p[None, :]
    add = rng.standard_normal((m, T))
    dom = 0.5 * rng.standard_normal((m, T))
    sire = np.array([0, 17, 2500, 4999, 4999, 1234, 77, 3001, 4242, 9], np.int32)
    dam = np.array([4999, 17, 2501, 0, 4999, 4321, 78, 3001, 13, 4998], np.int32)
    blocks = [O.recombfreq_to_ld_decay(O.genmap2recombfreq(cM[a:b], np.ones(b - a))) for a, b in ((0, 5556), (5556, m))]
    want, wseg = oracle_slab_c("AD", H, add, dom, None, sire, dam, blocks=blocks)
    return dict(m_c=m_c, cM=cM, H=H, add=add, dom=dom, sire=sire, dam=dam, P=P), want, wseg


def test_large_config_geometry_against_the_oracle(gms, large_geometry, mode):
    s, want, wseg = large_geometry
    eng = gms.CrossVarEngine(0)
    eng.set_mode(mode)
    eng.set_genmap(s["cM"], s["m_c"])
    eng.set_haplotypes(s["H"])
    out, nseg = eng.predict("AD", s["add"], s["dom"], None, s["sire"], s["dam"])
    st = eng.stats()
    assert st["last_mode"] == (2 if mode == "auto" else 0)
    assert np.array_equal(nseg, wseg)
    from tests.util import strict_rel_err
    assert rel_err(out, want) < TOL and strict_rel_err(out, want) < TOL
    # the same crosses inside a larger batch (more cross tiles, another plan) give the same numbers
    rng = np.random.default_rng(5)
    sire = np.concatenate([s["sire"], rng.integers(0, s["P"], 500).astype(np.int32)])
    dam = np.concatenate([s["dam"], rng.integers(0, s["P"], 500).astype(np.int32)])
    big, nbig = eng.predict("AD", s["add"], s["dom"], None, sire, dam)
    assert np.array_equal(nbig[:10], wseg) and rel_err(big[:10], want) < TOL
    eng.close()


# ---------------------------------------------------------------- map-structured fast path
@pytest.mark.parametrize("model", ["A", "AD", "DirDom"])
@pytest.mark.parametrize("T", [1, 2, 5, 6, 7, 11])
def test_map_scan_mode_matches_oracle_and_dense(gms, model, T):
    s = synth(300 + T, [200, 131, 77, 300], P=24, T=T)
    rng = np.random.default_rng(6)
    sire = rng.integers(0, 24, 90).astype(np.int32)
    dam = rng.integers(0, 24, 90).astype(np.int32)
    sire[:5] = dam[:5]
    dom = s["dom"] if model != "A" else None
    asub = s["asub"] if model == "DirDom" else None
    eng = gms.CrossVarEngine(0)
    eng.set_genmap(s["cM"], s["m_c"])
    eng.set_haplotypes(s["H"])
    eng.set_mode("fp64")
    dense, nseg_d = eng.predict(model, s["add"], dom, asub, sire, dam)
    eng.set_mode("map_scan")
    out, nseg = eng.predict(model, s["add"], dom, asub, sire, dam)
    want, wseg = oracle_slab(model, s["H"], s["R"], s["add"], dom, asub, sire, dam)
    assert np.array_equal(nseg, wseg) and np.array_equal(nseg_d, wseg)
    assert rel_err(out, want) < TOL
    assert rel_err(out, dense) < TOL
    eng.set_mode("fp64")
    again, _ = eng.predict(model, s["add"], dom, asub, sire, dam)
    assert np.array_equal(again, dense)                     # the FP64 path is bit-reproducible
    eng.close()


def test_map_scan_needs_a_sorted_map(gms):
    s = synth(4, [90, 60], P=4, T=2)
    eng = gms.CrossVarEngine(0)
    eng.set_recomb_matrix(s["R"], s["chrom"])
    eng.set_haplotypes(s["H"])
    with pytest.raises(gms.GmsError, match="sorted"):
        eng.set_mode("map_scan")
    cM = s["cM"].copy()
    cM[[3, 4]] = cM[[4, 3]]                               # unsorted map: dense still fine, scan refused
    eng.set_genmap(cM, s["m_c"])
    eng.set_haplotypes(s["H"])
    with pytest.raises(gms.GmsError, match="sorted"):
        eng.set_mode("map_scan")
    R = O.recombfreq_to_ld_decay(O.genmap2recombfreq(cM, s["chrom"]))
    out, nseg = eng.predict("AD", s["add"], s["dom"], None, [0, 1], [2, 3])
    want, wseg = oracle_slab("AD", s["H"], R, s["add"], s["dom"], None, [0, 1], [2, 3])
    assert np.array_equal(nseg, wseg) and rel_err(out, want) < TOL
    eng.close()


def test_map_scan_cassava_scale_equals_dense(gms):
    from genomicmateselectr_b200 import synth as S
    d = S.make("cassava", N=3000)
    eng = gms.CrossVarEngine(0)
    eng.set_genmap(d["cM"], d["m_c"])
    eng.set_haplotypes(d["H"])
    eng.set_mode("fp64")
    dense, nseg = eng.predict("AD", d["add"], d["dom"], None, d["sire"], d["dam"])
    eng.set_mode("map_scan")
    fast, nseg2 = eng.predict("AD", d["add"], d["dom"], None, d["sire"], d["dam"])
    assert np.array_equal(nseg, nseg2)
    assert rel_err(fast, dense) < TOL
    eng.close()


def test_predict_sharded_one_process(gms):
    """gms_predict_sharded: one host process, one handle per visible GPU (2 handles on the same GPU
    when only one is visible - the threads then just serialise on the device)."""
    import torch
    s = synth(21, [150, 140], P=10, T=3)
    rng = np.random.default_rng(1)
    sire = rng.integers(0, 10, 41).astype(np.int32)
    dam = rng.integers(0, 10, 41).astype(np.int32)
    ndev = max(1, torch.cuda.device_count())
    engines = [gms.CrossVarEngine(i % ndev) for i in range(max(2, min(ndev, 4)))]
    for e in engines:
        e.set_genmap(s["cM"], s["m_c"])
        e.set_haplotypes(s["H"])
    out, nseg = gms.CrossVarEngine.predict_sharded(engines, "AD", s["add"], s["dom"], None, sire, dam)
    ref, rseg = engines[0].predict("AD", s["add"], s["dom"], None, sire, dam)
    assert np.array_equal(nseg, rseg) and np.array_equal(out, ref)
    with pytest.raises(gms.GmsError, match="out of range"):
        bad = dam.copy(); bad[-1] = 99
        gms.CrossVarEngine.predict_sharded(engines, "AD", s["add"], s["dom"], None, sire, bad)
    for e in engines:
        e.close()


# ---------------------------------------------------------------- exact int8 tensor-core path (tcgen05)
@pytest.mark.parametrize("model,T", [("AD", 1), ("AD", 2), ("AD", 5), ("DirDom", 3), ("AD", 6), ("A", 4),
                                     ("AD", 7), ("AD", 10), ("DirDom", 13), ("AD", 24)])
def test_int8_tensor_mode_matches_oracle_and_dense(gms, model, T):
    s = synth(500 + T, [200, 131, 77, 300], P=24, T=T)
    rng = np.random.default_rng(8)
    sire = rng.integers(0, 24, 150).astype(np.int32)
    dam = rng.integers(0, 24, 150).astype(np.int32)
    sire[:5] = dam[:5]
    asub = s["asub"] if model == "DirDom" else None
    dom = s["dom"] if model != "A" else None
    eng = gms.CrossVarEngine(0)
    eng.set_genmap(s["cM"], s["m_c"])
    eng.set_haplotypes(s["H"])
    eng.set_mode("fp64")
    dense, nseg_d = eng.predict(model, s["add"], dom, asub, sire, dam)
    eng.set_mode("int8_tensor")
    out, nseg = eng.predict(model, s["add"], dom, asub, sire, dam)
    assert eng.stats()["last_mode"] == 2
    want, wseg = oracle_slab(model, s["H"], s["R"], s["add"], dom, asub, sire, dam)
    assert np.array_equal(nseg, wseg)
    assert rel_err(dense, want) < TOL
    assert rel_err(out, want) < TOL
    assert rel_err(out, dense) < TOL
    eng.close()


def test_int8_tensor_general_matrix_and_many_tiles(gms):
    """A non-map (one block) matrix and > 128 crosses x several 64-row tiles."""
    rng = np.random.default_rng(13)
    m, P, T = 330, 30, 4
    B = rng.standard_normal((m, 40))
    R = B @ B.T / 40
    H = (rng.random((2 * P, m)) < 0.4).astype(np.uint8)
    add, dom = rng.standard_normal((m, T)), rng.standard_normal((m, T)) * np.exp(rng.normal(0, 2, (m, 1)))
    sire = rng.integers(0, P, 300).astype(np.int32)
    dam = rng.integers(0, P, 300).astype(np.int32)
    eng = gms.CrossVarEngine(0)
    eng.set_recomb_matrix(R, np.ones(m, int))
    eng.set_haplotypes(H)
    eng.set_mode("int8_tensor")
    out, nseg = eng.predict("AD", add, dom, None, sire, dam)
    want, wseg = oracle_slab("AD", H, R, add, dom, None, sire, dam)
    assert np.array_equal(nseg, wseg) and rel_err(out, want) < TOL
    eng.close()


def test_trait_count_limits(gms):
    s = synth(31, [150, 90], P=6, T=32)
    eng = gms.CrossVarEngine(0)
    eng.set_genmap(s["cM"], s["m_c"])
    eng.set_haplotypes(s["H"])
    sire, dam = np.array([0, 1, 2], np.int32), np.array([3, 1, 5], np.int32)
    out, nseg = eng.predict("AD", s["add"], s["dom"], None, sire, dam)            # T = 32: 2 units per tile
    want, wseg = oracle_slab("AD", s["H"], s["R"], s["add"], s["dom"], None, sire, dam)
    assert out.shape == (3, 528, 2) and np.array_equal(nseg, wseg) and rel_err(out, want) < TOL
    assert eng.stats()["last_mode"] == 2                                           # auto -> exact int8 tensor path
    eng.set_mode("fp64")
    again, _ = eng.predict("AD", s["add"], s["dom"], None, sire, dam)
    assert eng.stats()["last_mode"] == 0 and rel_err(again, want) < TOL
    big = np.zeros((s["m"], 33))
    with pytest.raises(gms.GmsError, match="1..32"):
        eng.predict("A", big, None, None, sire, dam)
    eng.close()
