"""The R .Call shim (r-package/src/gms_rglue.c) EXECUTED end to end on the GPU box, without R: it is compiled with a
minimal stand-in for R's C API (r-package/tests/fakeR/fake_r.c: typed vectors, attributes, external pointers with
finalizers, the routine registration table) and driven by r-package/tests/fakeR/run_shim.c the way
r-package/R/predCrossVars.R and predictCrosses.R drive it: gmsr_create -> gmsr_set_recomb_blocks / gmsr_set_genmap ->
gmsr_set_haplotypes -> gmsr_predict -> gmsr_tidy -> gmsr_close, all through the routines registered by
R_init_genomicMateSelectRb200. Checked: the vignette numbers on the bundled data, the oracle on a synthetic map
problem, PROTECT balance, error propagation (an R condition = exit status 3 with the library's message)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import predcrossvar_oracle as O
from tests.util import oracle_slab, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    from genomicmateselectr_b200 import _capi
    _capi.lib()
    exe = str(tmp_path_factory.mktemp("rshim") / "run_shim")
    libdir = os.path.dirname(_capi.LIB_PATH)
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-I", os.path.join(ROOT, "r-package", "tests", "rstub"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "r-package", "src", "gms_rglue.c"),
           os.path.join(ROOT, "r-package", "tests", "fakeR", "fake_r.c"),
           os.path.join(ROOT, "r-package", "tests", "fakeR", "run_shim.c"),
           "-o", exe, "-L", libdir, "-lgms_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def run_shim(exe, tmp_path, model, m_c, H, add, dom, asub, sire, dam, blocks=None, cM=None, siwts=None, stdselint=2.67):
    """H: (2P) x m; written as an R integer matrix (column-major). Returns (out[N, npairs, ncomp], nseg, tidy, rc, stderr)."""
    m, P2 = H.shape[1], H.shape[0]
    T, N = add.shape[1], len(sire)
    prob, res = str(tmp_path / "problem.bin"), str(tmp_path / "result.bin")
    with open(prob, "wb") as f:
        f.write(struct.pack("<8i", model, len(m_c), m, P2 // 2, T, N, 1 if cM is not None else 0, 1 if siwts is not None else 0))
        f.write(struct.pack("<d", stdselint))
        f.write(np.asarray(m_c, "<i4").tobytes())
        if cM is not None:
            f.write(np.asarray(cM, "<f8").tobytes())
        else:
            for b in blocks:
                f.write(np.asfortranarray(b, dtype="<f8").tobytes(order="F"))
        f.write(np.asfortranarray(H.astype("<i4")).tobytes(order="F"))
        for e in (add, dom, asub):
            if e is not None:
                f.write(np.asfortranarray(e, dtype="<f8").tobytes(order="F"))
        f.write(np.asarray(sire, "<i4").tobytes())
        f.write(np.asarray(dam, "<i4").tobytes())
        if siwts is not None:
            f.write(np.asarray(siwts, "<f8").tobytes())
    r = subprocess.run([exe, prob, res], capture_output=True, text=True, timeout=300)
    if r.returncode != 0:
        return None, None, None, r.returncode, r.stderr
    raw = open(res, "rb").read()
    ncomp, npairs, n, nof, nslots = struct.unpack_from("<5i", raw, 0)
    off = 20
    out = np.frombuffer(raw, "<f8", ncomp * npairs * n, off).reshape(n, npairs, ncomp)      # R array dim c(ncomp, npairs, N)
    off += out.nbytes
    nseg = np.frombuffer(raw, "<i4", n, off)
    off += nseg.nbytes
    tidy = np.frombuffer(raw, "<f8", 4 * nslots * nof * n, off).reshape(n, nof, nslots, 4)  # dim c(4, nslots, nof, N)
    off += tidy.nbytes
    depth = struct.unpack_from("<i", raw, off)[0]
    assert depth == 0, "PROTECT / UNPROTECT imbalance in the shim"
    return out, nseg, tidy, 0, r.stderr


def test_shim_bundled_AD_matches_the_vignette(shim, bundled, tmp_path):
    b = bundled
    idx = []
    for p in b.PARENTS:
        idx += [b.hap_rownames.index(f"{p}_HapA"), b.hap_rownames.index(f"{p}_HapB")]
    H = b.haploMat[idx]
    sire, dam = zip(*[(i, j) for i in range(5) for j in range(i, 5)])
    blocks = [b.recombFreqMat[:50, :50], b.recombFreqMat[50:, 50:]]
    out, nseg, tidy, rc, err = run_shim(shim, tmp_path, 1, [50, 50], H, b.effAD_add, b.effAD_dom, None, sire, dam,
                                        blocks=blocks, siwts=[0.75, 0.25])
    assert rc == 0, err
    want, wseg = oracle_slab("AD", H, b.recombFreqMat, b.effAD_add, b.effAD_dom, None, sire, dam)
    assert np.array_equal(nseg, wseg) and rel_err(out, want) < 1e-9
    x = 14                                                            # 146 x 146
    assert [round(v) for v in out[x, :, 0]] == [1232, 1976, -820]     # non_additive_models.html:624-626
    assert [round(v, 2) for v in out[x, :, 1]] == [9.97, 51.27, 9.04]
    # tidy: predOf BV / TGV, slot 0 = SELIND: 509 / 521 (non_additive_models.html:600-601)
    assert round(tidy[x, 0, 0, 1]) == 509 and round(tidy[x, 1, 0, 1]) == 521
    assert np.allclose(tidy[:, :, :, 2], np.sqrt(tidy[:, :, :, 1]))
    assert np.allclose(tidy[:, :, :, 3], tidy[:, :, :, 0] + 2.67 * tidy[:, :, :, 2])


def test_shim_bundled_A_matches_the_vignette(shim, bundled, tmp_path):
    b = bundled
    idx = []
    for p in b.PARENTS:
        idx += [b.hap_rownames.index(f"{p}_HapA"), b.hap_rownames.index(f"{p}_HapB")]
    H = b.haploMat[idx]
    sire, dam = zip(*[(i, j) for i in range(5) for j in range(i, 5)])
    blocks = [b.recombFreqMat[:50, :50], b.recombFreqMat[50:, 50:]]
    out, nseg, tidy, rc, err = run_shim(shim, tmp_path, 0, [50, 50], H, b.effA, None, None, sire, dam, blocks=blocks,
                                        siwts=[0.75, 0.25])
    assert rc == 0, err
    assert list(nseg[:4]) == [22, 46, 48, 53]
    assert [round(v) for v in out[0, :, 0]] == [707, 1519, -350]      # start_here.html:385
    assert [round(v) for v in tidy[:4, 0, 0, 1]] == [361, 454, 458, 536]     # SELIND variances, start_here.html:333-342


def test_shim_map_path_dirdom_against_the_oracle(shim, tmp_path):
    rng = np.random.default_rng(5)
    m_c = [130, 77]
    m = sum(m_c)
    cM = np.concatenate([np.sort(rng.uniform(0, 100, k)) for k in m_c])
    chrom = np.repeat([1, 2], m_c)
    H = (rng.random((12, m)) < 0.5).astype(np.uint8)
    add, dom = rng.standard_normal((m, 2)), rng.standard_normal((m, 2))
    asub = add + 0.25 * dom
    sire, dam = [0, 1, 2, 5], [3, 1, 4, 0]
    out, nseg, tidy, rc, err = run_shim(shim, tmp_path, 2, m_c, H, add, dom, asub, sire, dam, cM=cM)
    assert rc == 0, err
    R = O.recombfreq_to_ld_decay(O.genmap2recombfreq(cM, chrom))
    want, wseg = oracle_slab("DirDom", H, R, add, dom, asub, sire, dam)
    assert np.array_equal(nseg, wseg) and rel_err(out, want) < 1e-9
    assert tidy.shape == (4, 2, 2, 4)
    assert np.allclose(tidy[:, 0, :, 1], want[:, :2, 2], rtol=1e-9)                   # VarBV
    assert np.allclose(tidy[:, 1, :, 1], want[:, :2, 0] + want[:, :2, 1], rtol=1e-9)  # VarTGV = VarA + VarD


def test_shim_turns_library_errors_into_r_conditions(shim, tmp_path):
    rng = np.random.default_rng(6)
    m_c = [40]
    cM = np.sort(rng.uniform(0, 50, 40))
    H = (rng.random((4, 40)) < 0.5).astype(np.uint8)
    H[1, 3] = 2                                                       # not a haplotype
    add = rng.standard_normal((40, 1))
    _, _, _, rc, err = run_shim(shim, tmp_path, 0, m_c, H, add, None, None, [0], [1], cM=cM)
    assert rc == 3 and "gms_b200" in err and "outside {0,1}" in err
    H[1, 3] = 1
    _, _, _, rc, err = run_shim(shim, tmp_path, 0, m_c, H, add, None, None, [0], [7], cM=cM)
    assert rc == 3 and "out of range" in err
