"""Times the per-cross int8 kernel under the test / profiling hooks, one engine per configuration in ONE process:
    python tools/i8_probe.py N "SORT,SKIP,DEBUG" ["SORT,SKIP,DEBUG" ...]
SORT: GMS_I8_SORT (api.cu); SKIP: ignored since the row-skipping epilogue was dropped (kept so that old command lines still parse); DEBUG: GMS_I8_DEBUG, only honoured by a -DGMS_PROFILE build selected with
GMS_B200_LIB (bits: 1 no epilogue, 2 no MMAs, 4 no second MMA, 8 no xi copy, 16 no digit copy, 32 drain but no arithmetic;
results are garbage then). Prints the kernel time (CUDA events inside the library, best of 3) per configuration."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomicmateselectr_b200 import CrossVarEngine, synth, _capi

N = int(sys.argv[1])
workload = os.environ.get("PROBE_WORKLOAD", "cassava")
d = synth.make(workload, N=N)
print("lib", _capi.LIB_PATH, "workload", workload, "N", N, flush=True)
for cfg in sys.argv[2:]:
    so, sk, dbg = cfg.split(",")
    os.environ["GMS_I8_SORT"], os.environ["GMS_I8_SKIP"], os.environ["GMS_I8_DEBUG"] = so, sk, dbg
    eng = CrossVarEngine(0)
    eng.set_cache(False)
    eng.set_genmap(d["cM"], d["m_c"])
    eng.set_haplotypes(d["H"])
    eng.set_mode("int8_tensor")
    if int(dbg):
        eng.set_int8_tolerance(0.0)          # garbage results must not trigger the FP64 re-evaluation
    eng.stage("AD", d["add"], d["dom"], None, d["sire"], d["dam"])
    ks, tot, par = [], [], []
    for _ in range(3):
        eng.compute(sync=True)
        st = eng.stats()
        ks.append(st["last_ms_cross_kernel"]); tot.append(st["last_ms_total"]); par.append(st["last_ms_parent_stage"])
    out, nseg = eng.fetch()
    print(f"sort={so} skip={sk} debug={dbg:>3s}: cross kernel {min(ks):8.3f} ms  parents+prep {min(par):7.3f} ms  total {min(tot):8.3f} ms  "
          f"checksum {float(np.abs(out).sum()):.6e}  fallback {eng.stats()['last_fallback_units']}", flush=True)
    eng.close()
