"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol declared in
include/gms_b200.h, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gms_b200.h")).read()
    return sorted(set(re.findall(r"^GMS_API\s+[\w\s\*]+?\b(gms_\w+)\s*\(", src, flags=re.M)))


def test_header_and_binding_agree():
    from genomicmateselectr_b200 import _capi
    assert declared_symbols() == _capi.EXPORTS


def test_library_exports_every_declared_symbol():
    from genomicmateselectr_b200 import _capi
    L = _capi.lib()
    for name in declared_symbols():
        assert hasattr(L, name), name
    assert b"sm_100a" in L.gms_version()


def test_trait_pair_order():
    from genomicmateselectr_b200 import trait_pairs
    a, b = trait_pairs(3)
    assert list(zip(a, b)) == [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]   # R/predCrossVar.R:128-138
    a, b = trait_pairs(1)
    assert list(zip(a, b)) == [(0, 0)]


def test_selind_is_host_side_and_matches_oracle(bundled):
    from genomicmateselectr_b200 import CrossVarEngine
    from oracle import predcrossvar_oracle as O
    rng = np.random.default_rng(0)
    out = rng.standard_normal((4, 6, 2))
    w = np.array([0.5, 0.3, 0.2])
    si = CrossVarEngine.selind("AD", out, w)
    traits = ["a", "b", "c"]
    pairs = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]
    for x in range(4):
        for c in range(2):
            rows = [(traits[i], traits[j], out[x, k, c]) for k, (i, j) in enumerate(pairs)]
            assert si[x, c] == pytest.approx(O.selind_var(rows, traits, dict(zip(traits, w))), rel=1e-13)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from genomicmateselectr_b200 import CrossVarEngine, GmsError
    with pytest.raises(GmsError, match="no CUDA device"):
        CrossVarEngine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "genomicmateselectr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("no cpu fallback", ""), f
