"""Generate tests/golden/bundled.npz from the reference's bundled data/*.rda.

Run ONCE in the build container (where /root/reference is mounted read-only):

    python tests/golden/make_fixtures.py

The output is committed; nothing at test/bench time reads /root/reference.
Objects and their provenance:
  haploMat        data/haploMat.rda       int[600 x 100], rows <GID>_HapA/_HapB
  recombFreqMat   data/recombFreqMat.rda  dbl[100 x 100] (already 1-2c; built by an
                                          older genmap2recombfreq without the /100)
  genmap          data/genmap.rda         named dbl[100] (cM; names "<chr>_<id>")
  doseMat         data/doseMat.rda        int[300 x 100] (= HapA + HapB)
  snpeffsA/AD/DirDom  data/snpeffs*.rda   tibbles with list-columns of dbl[100 x 1]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import rdx2  # noqa: E402

REF = "/root/reference/data"
OUT = os.path.join(os.path.dirname(__file__), "bundled.npz")


def effcol(tib, col):
    names = tib.names()
    lst = tib.value[names.index(col)].value
    mats = []
    snps = None
    for e in lst:
        a, rn, _ = rdx2.matrix(e)
        assert a.shape[1] == 1
        mats.append(a[:, 0])
        snps = rn
    return np.stack(mats, axis=1), snps  # m x T


def main():
    out = {}
    hap, hrn, hcn = rdx2.matrix(rdx2.load_rda(f"{REF}/haploMat.rda")["haploMat"])
    assert set(np.unique(hap)) <= {0, 1}
    out["haploMat"] = hap.astype(np.uint8)
    out["hap_rownames"] = np.array(hrn)
    out["snp_names"] = np.array(hcn)
    R, rrn, rcn = rdx2.matrix(rdx2.load_rda(f"{REF}/recombFreqMat.rda")["recombFreqMat"])
    assert rrn == hcn and rcn == hcn
    out["recombFreqMat"] = R
    gm = rdx2.load_rda(f"{REF}/genmap.rda")["genmap"]
    assert gm.names() == hcn
    out["genmap"] = np.asarray(gm.value)
    dose, drn, dcn = rdx2.matrix(rdx2.load_rda(f"{REF}/doseMat.rda")["doseMat"])
    assert dcn == hcn
    out["doseMat"] = dose.astype(np.uint8)
    out["dose_rownames"] = np.array(drn)
    for name, cols in (("snpeffsA", ["allelesubsnpeff"]),
                       ("snpeffsAD", ["allelesubsnpeff", "domdevsnpeff"]),
                       ("snpeffsDirDom", ["allelesubsnpeff", "domdevsnpeff"])):
        tib = rdx2.load_rda(f"{REF}/{name}.rda")[name]
        traits = list(tib.value[tib.names().index("Trait")].value)
        out[f"{name}_traits"] = np.array(traits)
        for c in cols:
            eff, snps = effcol(tib, c)
            assert snps == hcn
            out[f"{name}_{c}"] = eff
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
