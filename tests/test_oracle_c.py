"""The oracle's C restatement (used for the bounded CPU baseline) against the NumPy oracle, which is
itself pinned to the reference's vignette outputs."""
import numpy as np
import pytest

from oracle import crossvar_ref
from oracle import predcrossvar_oracle as O
from tests.util import oracle_slab


def test_c_oracle_matches_numpy_oracle_on_bundled_data(bundled):
    b = bundled
    idx = []
    for p in b.PARENTS:
        idx += [b.hap_rownames.index(f"{p}_HapA"), b.hap_rownames.index(f"{p}_HapB")]
    H = b.haploMat[idx]
    prob = crossvar_ref.RefProblem(H, b.effAD_add, b.effAD_dom, R=b.recombFreqMat)
    sire, dam = zip(*[(i, j) for i in range(5) for j in range(i, 5)])
    want, wseg = oracle_slab("AD", H, b.recombFreqMat, b.effAD_add, b.effAD_dom, None, sire, dam)
    for x, (s, d) in enumerate(zip(sire, dam)):
        out, nseg = prob.one_cross(s, d, "AD", nthreads=2)
        assert nseg == wseg[x]
        assert np.allclose(out, want[x], rtol=1e-12, atol=0)
    out, nseg = prob.one_cross(4, 4, "AD")          # 146 x 146, non_additive_models.html:624-629
    assert nseg == 34 and [round(v) for v in out[:, 0]] == [1232, 1976, -820]


@pytest.mark.parametrize("model", ["A", "AD"])
def test_c_oracle_block_form_equals_dense_form(model):
    rng = np.random.default_rng(1)
    m_c = [90, 61, 40]
    m = sum(m_c)
    cM = np.concatenate([np.sort(rng.uniform(0, 100, k)) for k in m_c])
    chrom = np.repeat([1, 2, 3], m_c)
    R = O.recombfreq_to_ld_decay(O.genmap2recombfreq(cM, chrom))
    offs = np.concatenate([[0], np.cumsum(m_c)])
    blocks = [R[a:b, a:b] for a, b in zip(offs[:-1], offs[1:])]
    H = (rng.random((12, m)) < 0.5).astype(np.uint8)
    add, dom = rng.standard_normal((m, 3)), rng.standard_normal((m, 3))
    dense = crossvar_ref.RefProblem(H, add, dom, R=R)
    blk = crossvar_ref.RefProblem(H, add, dom, blocks=blocks)
    want, wseg = oracle_slab(model, H, R, add, dom if model == "AD" else None, None, [0, 1, 5], [3, 1, 2])
    for x, (s, d) in enumerate([(0, 3), (1, 1), (5, 2)]):
        o1, n1 = dense.one_cross(s, d, model)
        o2, n2 = blk.one_cross(s, d, model, nthreads=3)
        assert n1 == n2 == wseg[x]
        assert np.allclose(o1, want[x], rtol=1e-12, atol=0) and np.allclose(o2, want[x], rtol=1e-12, atol=0)


@pytest.mark.parametrize("model", ["A", "AD", "DirDom"])
def test_c_oracle_map_form_equals_numpy_oracle(model):
    """recombFreqMat given as the genetic map (entries of 1-2*genmap2recombfreq() evaluated on the fly): what the
    full-size parity tests use where the m x m matrix would not fit."""
    from tests.util import oracle_slab_c
    rng = np.random.default_rng(4)
    m_c = [120, 75, 33]
    m = sum(m_c)
    cM = np.concatenate([np.sort(rng.uniform(0, 100, k)) for k in m_c])
    chrom = np.repeat([1, 2, 3], m_c)
    R = O.recombfreq_to_ld_decay(O.genmap2recombfreq(cM, chrom))
    H = (rng.random((12, m)) < 0.5).astype(np.uint8)
    add, dom = rng.standard_normal((m, 3)), rng.standard_normal((m, 3))
    asub = add + 0.2 * dom
    sire, dam = [0, 1, 5, 2], [3, 1, 2, 2]
    want, wseg = oracle_slab(model, H, R, add, dom if model != "A" else None, asub if model == "DirDom" else None, sire, dam)
    got, nseg = oracle_slab_c(model, H, add, dom, asub, sire, dam, cM=cM, chrom=chrom, nthreads=2)
    assert np.array_equal(nseg, wseg)
    assert np.allclose(got, want, rtol=1e-12, atol=0)
