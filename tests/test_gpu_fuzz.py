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
