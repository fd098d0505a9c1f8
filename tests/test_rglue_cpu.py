"""The R .Call shim cannot be built here (no R). It is at least compiled (syntax + types against the
real include/gms_b200.h and stub R headers) so that it cannot rot silently."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rglue_compiles_against_the_c_abi():
    cmd = ["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only",
           "-I", os.path.join(ROOT, "r-package", "tests", "rstub"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "r-package", "src", "gms_rglue.c")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_rglue_registers_every_entry_point_it_defines():
    import re
    src = open(os.path.join(ROOT, "r-package", "src", "gms_rglue.c")).read()
    defined = set(re.findall(r"^SEXP (gmsr_\w+)\(", src, flags=re.M))
    registered = set(re.findall(r'\{"(gmsr_\w+)"', src))
    assert defined == registered
    rsrc = open(os.path.join(ROOT, "r-package", "R", "predCrossVars.R")).read()
    assert set(re.findall(r'\.Call\("(gmsr_\w+)"', rsrc)) <= registered


def test_r_sources_only_call_registered_routines():
    import re
    src = open(os.path.join(ROOT, "r-package", "src", "gms_rglue.c")).read()
    registered = set(re.findall(r'\{"(gmsr_\w+)"', src))
    for name in ("predCrossVars.R", "predictCrosses.R"):
        rsrc = open(os.path.join(ROOT, "r-package", "R", name)).read()
        assert set(re.findall(r'\.Call\("(gmsr_\w+)"', rsrc)) <= registered, name


def test_shim_links_against_the_fake_r_api_and_reports_errors(tmp_path):
    """The shim + the stand-in for R's C API (r-package/tests/fakeR) + the driver link against libgms_b200.so; without a
    GPU the first routine fails and the shim must turn that into an R condition (exit status 3 + the library's text).
    The full run is tests/test_gpu_rshim.py."""
    import struct
    import torch
    from genomicmateselectr_b200 import _capi
    _capi.lib()
    libdir = os.path.dirname(_capi.LIB_PATH)
    exe = str(tmp_path / "run_shim")
    cmd = ["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "r-package", "tests", "rstub"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "r-package", "src", "gms_rglue.c"),
           os.path.join(ROOT, "r-package", "tests", "fakeR", "fake_r.c"),
           os.path.join(ROOT, "r-package", "tests", "fakeR", "run_shim.c"),
           "-o", exe, "-L", libdir, "-lgms_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if torch.cuda.is_available():
        return
    prob = tmp_path / "p.bin"
    prob.write_bytes(struct.pack("<8i", 0, 1, 4, 1, 1, 1, 1, 0) + struct.pack("<d", 2.67) + struct.pack("<i", 4) + b"\0" * 512)
    r = subprocess.run([exe, str(prob), str(tmp_path / "r.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr
