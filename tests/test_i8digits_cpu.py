"""Known-answer test of the fixed-point arithmetic the int8 tensor path runs on the device (csrc/i8digits.cuh is shared
between the kernels and this host build): balanced radix-256 digit split, exact 64-bit recombination of the digit sums for
up to 32767 same-sign terms, and the integer-only FP64 conversion of the epilogue against exact arithmetic."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("ndig", [6, 7])
def test_i8digits_known_answers(tmp_path, ndig):
    exe = str(tmp_path / "i8digits_kat")
    subprocess.run(["g++", "-O2", "-std=c++17", f"-DGMS_I8_DIG={ndig}", "-o", exe,
                    os.path.join(ROOT, "tests", "csrc", "i8digits_kat.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout + r.stderr


def test_i8_block_ownership(tmp_path):
    """csrc/common.cuh: the round-robin deal of a chromosome's 64-blocks to its 64-row tiles covers every unordered pair once."""
    exe = str(tmp_path / "i8blocks_kat")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", cuda_inc, "-o", exe,
                    os.path.join(ROOT, "tests", "csrc", "i8blocks_kat.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout + r.stderr
