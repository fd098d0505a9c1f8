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
