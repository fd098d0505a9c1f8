"""Edge-shape fuzz of the CUDA path against the oracle, in every evaluation mode: chromosome sizes around the
64 / 128 tile boundaries (1, 63, 64, 65, 127, 128, 129 ...), single-SNP chromosomes, one parent, cross counts around
the 128-cross and 80/T-unit tile boundaries."""
import numpy as np
import pytest

from oracle import predcrossvar_oracle as O
from tests.util import oracle_slab, rel_err

pytestmark = pytest.mark.gpu

CASES = [
    # (m_c, P, N, T, model)
    ([1], 2, 3, 1, "AD"),
    ([1, 1, 1], 3, 5, 2, "AD"),
    ([63], 2, 1, 3, "AD"),
    ([64], 3, 127, 1, "DirDom"),
    ([65], 1, 1, 5, "AD"),
    ([127, 128, 129], 5, 128, 2, "AD"),
    ([200, 1, 64], 4, 129, 6, "AD"),
    ([129, 257], 6, 300, 7, "AD"),
    ([300], 9, 17, 9, "A"),
    ([31, 33, 95, 97], 7, 81, 4, "DirDom"),
    ([513], 3, 16, 5, "AD"),
    ([70, 70, 70, 70, 70, 70], 8, 260, 3, "A"),
]


def make(seed, m_c, P, T):
    rng = np.random.default_rng(seed)
    m_c = np.asarray(m_c)
    m = int(m_c.sum())
    cM = np.concatenate([np.sort(rng.uniform(0, 90, k)) for k in m_c])
    chrom = np.repeat(np.arange(1, len(m_c) + 1), m_c)
    H = (rng.random((2 * P, m)) < rng.uniform(0.2, 0.8, m)).astype(np.uint8)
    add = rng.standard_normal((m, T))
    dom = rng.standard_normal((m, T)) * np.exp(rng.normal(0, 1.5, (m, 1)))      # wide dynamic range of effects
    asub = add + 0.3 * dom
    R = O.recombfreq_to_ld_decay(O.genmap2recombfreq(cM, chrom))
    return m_c, cM, chrom, H, add, dom, asub, R


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("mode", ["auto", "fp64", "map_scan"])
def test_edge_shapes(case, mode):
    import genomicmateselectr_b200 as g
    m_c, P, N, T, model = CASES[case]
    m_c, cM, chrom, H, add, dom, asub, R = make(1000 + case, m_c, P, T)
    rng = np.random.default_rng(case)
    sire = rng.integers(0, P, N).astype(np.int32)
    dam = rng.integers(0, P, N).astype(np.int32)
    d = dom if model != "A" else None
    a = asub if model == "DirDom" else None
    with g.CrossVarEngine(0) as eng:
        eng.set_genmap(cM, m_c)
        eng.set_haplotypes(H)
        eng.set_mode(mode)
        out, nseg = eng.predict(model, add, d, a, sire, dam)
    want, wseg = oracle_slab(model, H, R, add, d, a, sire, dam)
    assert np.array_equal(nseg, wseg)
    assert rel_err(out, want) < 1e-9


def test_one_engine_many_calls():
    """State handling of a long-lived handle (the CV driver's usage): changing model, trait count, cross
    count, mode, chromosome shard and haplotypes between calls on ONE engine; every answer against the oracle."""
    import genomicmateselectr_b200 as g
    m_c, cM, chrom, H, add, dom, asub, R = make(5, [150, 90, 33], 12, 9)
    rng = np.random.default_rng(0)
    eng = g.CrossVarEngine(0)
    eng.set_genmap(cM, m_c)
    eng.set_haplotypes(H)
    plan = [("AD", 5, 200, "auto"), ("A", 2, 7, "fp64"), ("DirDom", 9, 130, "auto"), ("AD", 1, 1, "map_scan"),
            ("AD", 6, 300, "int8_tensor"), ("AD", 3, 0, "auto"), ("DirDom", 4, 50, "fp64"), ("A", 9, 64, "map_scan")]
    for model, T, N, mode in plan:
        sire = rng.integers(0, 12, N).astype(np.int32)
        dam = rng.integers(0, 12, N).astype(np.int32)
        d = dom[:, :T] if model != "A" else None
        a = asub[:, :T] if model == "DirDom" else None
        eng.set_mode(mode)
        out, nseg = eng.predict(model, add[:, :T], d, a, sire, dam)
        want, wseg = oracle_slab(model, H, R, add[:, :T], d, a, sire, dam)
        assert np.array_equal(nseg, wseg), (model, T, N, mode)
        if N:
            assert rel_err(out, want) < 1e-9, (model, T, N, mode)
        gebv, mbv, _ = eng.cross_means(add[:, :T], None, sire, dam)            # interleaved "next row" call
        assert np.allclose(gebv, (H[0::2].astype(float) + H[1::2]) @ add[:, :T], rtol=1e-12)
    # new haplotypes (fewer parents) on the same map, then a chromosome shard
    H2 = H[:8]
    eng.set_haplotypes(H2)
    eng.set_mode("auto")
    sire, dam = np.array([0, 1, 2, 3], np.int32), np.array([3, 3, 0, 2], np.int32)
    out, nseg = eng.predict("AD", add[:, :4], dom[:, :4], None, sire, dam)
    want, wseg = oracle_slab("AD", H2, R, add[:, :4], dom[:, :4], None, sire, dam)
    assert np.array_equal(nseg, wseg) and rel_err(out, want) < 1e-9
    acc, accn = np.zeros_like(out), np.zeros_like(nseg)
    for c0, c1 in ((0, 2), (2, 3)):
        eng.set_chromosome_shard(c0, c1)
        o, n = eng.predict("AD", add[:, :4], dom[:, :4], None, sire, dam)
        acc += o
        accn += n
    assert np.array_equal(accn, wseg) and rel_err(acc, want) < 1e-9
    with pytest.raises(g.GmsError, match="out of range"):
        eng.predict("A", add[:, :1], None, None, [9], [0])                      # parent 9 no longer exists
    eng.close()


@pytest.mark.parametrize("env", [
    {"GMS_I8_CG2": "0"},                                            # one CTA per 128 crosses
    {"GMS_I8_CG2": "1"},                                            # the default: tcgen05 cta_group::2 pairs (M = 256)
], ids=["plain", "cta_group2"])
@pytest.mark.parametrize("T", [1, 3, 5, 6, 7])
def test_int8_kernel_variants(monkeypatch, env, T):
    """Both launch forms of the int8 tcgen05 kernel (environment switch read at gms_create) against the oracle: the default
    (cta_group::2 pairs, each CTA staging half of the digit operand) and the single-CTA form."""
    import genomicmateselectr_b200 as g
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    m_c, cM, chrom, H, add, dom, asub, R = make(9 + T, [200, 131, 77], 10, T)
    rng = np.random.default_rng(2)
    sire = rng.integers(0, 10, 300).astype(np.int32)
    dam = rng.integers(0, 10, 300).astype(np.int32)
    with g.CrossVarEngine(0) as eng:
        eng.set_genmap(cM, m_c)
        eng.set_haplotypes(H)
        eng.set_mode("int8_tensor")
        out, nseg = eng.predict("DirDom", add, dom, asub, sire, dam)
    want, wseg = oracle_slab("DirDom", H, R, add, dom, asub, sire, dam)
    assert np.array_equal(nseg, wseg) and rel_err(out, want) < 1e-9


def test_int8_cross_chunking(monkeypatch):
    """The per-cross int8 pass runs in chunks of crosses when the xi mask image would be too large
    (GMS_I8_CHUNK is a test hook that forces small chunks): full chunks + a ragged tail == one pass."""
    import genomicmateselectr_b200 as g
    m_c, cM, chrom, H, add, dom, asub, R = make(12, [150, 100], 9, 3)
    rng = np.random.default_rng(3)
    sire = rng.integers(0, 9, 700).astype(np.int32)
    dam = rng.integers(0, 9, 700).astype(np.int32)
    want, wseg = oracle_slab("AD", H, R, add, dom, None, sire, dam)
    for chunk in ("256", "512", "100000"):
        monkeypatch.setenv("GMS_I8_CHUNK", chunk)
        with g.CrossVarEngine(0) as eng:
            eng.set_genmap(cM, m_c)
            eng.set_haplotypes(H)
            eng.set_mode("int8_tensor")
            out, nseg = eng.predict("AD", add, dom, None, sire, dam)
            assert eng.stats()["last_mode"] == 2
        assert np.array_equal(nseg, wseg) and rel_err(out, want) < 1e-9, chunk


def test_int8_long_work_list():
    """One 40,000-SNP chromosome (625 64-row tiles) and enough crosses that the plan would use ONE range of tiles per CTA:
    the kernels keep their range of the work list in shared memory (512 entries), so the plan must cut it. A sample of
    crosses is checked against the oracle's C restatement with the matrix entries evaluated from the map (the 12.8 GB
    matrix itself is never formed), in the default mode and in FP64; the sample must also not depend on how many crosses
    are predicted with it (different plan)."""
    import genomicmateselectr_b200 as g
    from tests.util import oracle_slab_c
    rng = np.random.default_rng(11)
    m, P, T, N = 40000, 24, 2, 38400
    cM = np.sort(rng.uniform(0, 300, m))
    H = (rng.random((2 * P, m)) < rng.uniform(0.1, 0.9, m)).astype(np.uint8)
    add, dom = rng.standard_normal((m, T)), rng.standard_normal((m, T))
    sire = rng.integers(0, P, N).astype(np.int32)
    dam = rng.integers(0, P, N).astype(np.int32)
    pick = np.array([0, 299, 17000, N - 1])
    want, wseg = oracle_slab_c("AD", H, add, dom, None, sire[pick], dam[pick], cM=cM, chrom=np.ones(m, np.int32))
    with g.CrossVarEngine(0) as eng:
        eng.set_genmap(cM, np.array([m]))
        eng.set_haplotypes(H)
        eng.set_mode("int8_tensor")
        out, nseg = eng.predict("AD", add, dom, None, sire, dam)
        st = eng.stats()
        assert st["last_mode"] == 2 and st["last_fallback_overflow"] == 0
        few, nfew = eng.predict("AD", add, dom, None, sire[:300], dam[:300])
        eng.set_mode("fp64")
        o64, n64 = eng.predict("AD", add, dom, None, sire[pick], dam[pick])
    assert np.array_equal(nseg[pick], wseg) and np.array_equal(nfew, nseg[:300]) and np.array_equal(n64, wseg)
    assert rel_err(out[pick], want) < 1e-9
    assert rel_err(o64, want) < 1e-9
    assert rel_err(few, out[:300]) < 1e-12
