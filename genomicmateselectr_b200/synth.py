"""Deterministic synthetic inputs of the shapes BASELINE.json names.

cassava : P=1000 parents, m=33,000 SNPs over 18 chromosomes (12 x 1833 + 6 x 1834), T=5, N=50,000
large   : P=5000, m=100,000 (10 x 5556 + 8 x 5555), T=10, N=1,000,000
Map: per chromosome sorted U(0, 130 cM); allele frequency p_j ~ U(0.05, 0.95), haplotype bits
Bernoulli(p_j); effects add ~ N(0,1), dom ~ N(0, 0.25); crosses = distinct (sire <= dam) pairs."""
import numpy as np

CONFIGS = {
    "cassava": dict(P=1000, m_c=[1833] * 12 + [1834] * 6, T=5, N=50_000, seed=20261019),
    "large": dict(P=5000, m_c=[5556] * 10 + [5555] * 8, T=10, N=1_000_000, seed=20261119),
    "mini": dict(P=64, m_c=[300, 257, 130], T=3, N=500, seed=7),
}


def make(name="cassava", P=None, m_c=None, T=None, N=None, seed=None, haps=True):
    cfg = dict(CONFIGS[name])
    for k, v in dict(P=P, m_c=m_c, T=T, N=N, seed=seed).items():
        if v is not None:
            cfg[k] = v
    P, m_c, T, N, seed = cfg["P"], np.asarray(cfg["m_c"], np.int64), cfg["T"], cfg["N"], cfg["seed"]
    m = int(m_c.sum())
    rng = np.random.default_rng(seed)
    cM = np.concatenate([np.sort(rng.uniform(0.0, 130.0, int(k))) for k in m_c])
    chrom = np.repeat(np.arange(1, len(m_c) + 1), m_c)
    rng = np.random.default_rng(seed + 1)
    H = None
    if haps:
        p = rng.uniform(0.05, 0.95, m)
        H = np.empty((2 * P, m), np.uint8)
        step = max(1, (1 << 26) // m)
        for a in range(0, 2 * P, step):   # chunked to bound the float temporaries
            b = min(2 * P, a + step)
            H[a:b] = rng.random((b - a, m), dtype=np.float32) < p[None, :].astype(np.float32)
    rng = np.random.default_rng(seed + 2)
    add = rng.standard_normal((m, T))
    dom = 0.5 * rng.standard_normal((m, T))
    rng = np.random.default_rng(seed + 3)
    npairs = P * (P + 1) // 2
    if N > npairs:
        raise ValueError("more crosses than distinct pairs")
    pick = rng.choice(npairs, size=N, replace=False)
    # unrank (sire <= dam) pairs: pair index k -> row i of the upper triangle incl. diagonal
    i = np.floor((2 * P + 1 - np.sqrt((2 * P + 1) ** 2 - 8.0 * pick)) / 2).astype(np.int64)
    start = i * P - i * (i - 1) // 2
    fix = start > pick
    i[fix] -= 1
    start = i * P - i * (i - 1) // 2
    fix = pick - start >= P - i
    i[fix] += 1
    start = i * P - i * (i - 1) // 2
    j = i + (pick - start)
    assert np.all((j >= i) & (j < P) & (i >= 0))
    return dict(name=name, P=P, m=m, m_c=m_c, T=T, N=N, cM=cM, chrom=chrom, H=H, add=add, dom=dom,
                sire=i.astype(np.int32), dam=j.astype(np.int32), S2=int((m_c.astype(np.int64) ** 2).sum()))
