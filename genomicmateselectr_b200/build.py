"""Build libgms_b200.so (sm_100a) in-tree with nvcc. No JIT cache: the .so travels with the repo snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libgms_b200.so")
SOURCES = ["api.cu", "contract.cu", "kernels.cu", "scan.cu", "i8path.cu", "verify.cu"]
HEADERS = ["common.cuh", "kernels.cuh", "i8ptx.cuh", "i8digits.cuh", os.path.join("..", "..", "include", "gms_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines):
    """Another build of the same library for tools/ (never the product): lib/<name>/libgms_b200.so, loaded through GMS_B200_LIB.
    -DGMS_PROFILE compiles the profiling switches of csrc/i8path.cu in (GMS_I8_DEBUG then turns the MMAs, the copies or the
    epilogue arithmetic off - results are garbage); -DGMS_I8_DIG=6 selects six digit planes (default seven), -DGMS_I8_LATE=0 the FP64 burst right after the conversion,
    -DGMS_I8_TRIANGLE the lower block triangle instead of the even deal of 64-blocks, -DGMS_I8_TRAIT_OUTER the unit order trait > tile
    instead of chromosome > trait > tile."""
    out = os.path.join(LIBDIR, name)
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libgms_b200.so")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.run([_nvcc(), *NVCC_FLAGS, *defines, "-shared", "-o", lib, *srcs], check=True)
    return lib


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    for s in SOURCES:
        o = os.path.join(LIBDIR, s.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        objs.append(o)
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    if "--variants" in sys.argv:
        print(build_variant("profile", ["-DGMS_PROFILE"]))
        print(build_variant("dig6", ["-DGMS_I8_DIG=6"]))
        print(build_variant("traitouter", ["-DGMS_I8_TRAIT_OUTER"]))
        if "--all" in sys.argv:
            print(build_variant("early", ["-DGMS_I8_LATE=0"]))
        if "--all" in sys.argv:
            print(build_variant("tri", ["-DGMS_I8_TRIANGLE"]))
    else:
        print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
