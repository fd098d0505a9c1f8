import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Bundled:
    """The reference's bundled example data (tests/golden/bundled.npz; made by
    tests/golden/make_fixtures.py from /root/reference/data/*.rda)."""

    PARENTS = ["49", "153", "74", "228", "146"]  # set.seed(42) sample, docs/articles/start_here.html:281-288

    def __init__(self):
        z = np.load(os.path.join(ROOT, "tests", "golden", "bundled.npz"))
        self.haploMat = z["haploMat"]
        self.hap_rownames = [str(s) for s in z["hap_rownames"]]
        self.snp_names = [str(s) for s in z["snp_names"]]
        self.recombFreqMat = z["recombFreqMat"]
        self.genmap = z["genmap"]
        self.doseMat = z["doseMat"]
        self.dose_rownames = [str(s) for s in z["dose_rownames"]]
        self.traits = [str(s) for s in z["snpeffsA_traits"]]
        self.effA = z["snpeffsA_allelesubsnpeff"]
        self.effAD_add = z["snpeffsAD_allelesubsnpeff"]
        self.effAD_dom = z["snpeffsAD_domdevsnpeff"]
        self.effDD_asub = z["snpeffsDirDom_allelesubsnpeff"]
        self.effDD_domdev = z["snpeffsDirDom_domdevsnpeff"]
        self.chrom = np.array([int(s.split("_")[0]) for s in self.snp_names])

    def effdict(self, mat, traits=None):
        traits = traits or self.traits
        return {t: mat[:, i].copy() for i, t in enumerate(self.traits) if t in traits}

    def dirdom_effects(self, b=(-0.08215501, 0.05)):
        """Synthesise DirDom effect classes from bundled pieces (follows
        R/gmsFunctions.R:487-492): a = allelesub, d* = domdev, d = d* - b/m,
        alpha = a + d(q-p) with p = colMeans(dose)/2."""
        m = len(self.snp_names)
        p = self.doseMat.astype(np.float64).mean(axis=0) / 2.0
        q = 1.0 - p
        a = self.effDD_asub
        d = self.effDD_domdev - np.asarray(b)[None, :] / m
        alpha = a + d * (q - p)[:, None]
        return a, d, alpha


@pytest.fixture(scope="session")
def bundled():
    return Bundled()
