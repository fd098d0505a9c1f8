"""Debug: int8 path (verification off) against the FP64 path, per chromosome shard and per component."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import genomicmateselectr_b200 as g
from tests.test_gpu_parity import synth

def run(s, mode, model, add, dom, sire, dam, shard=None, tol=None):
    eng = g.CrossVarEngine(0)
    eng.set_genmap(s["cM"], s["m_c"])
    eng.set_haplotypes(s["H"])
    eng.set_mode(mode)
    if tol is not None:
        eng.set_int8_tolerance(tol)
    if shard is not None:
        eng.set_chromosome_shard(shard, shard + 1)
    out, nseg = eng.predict(model, add, dom, None, sire, dam)
    st = eng.stats()
    eng.close()
    return out, st

for m_c in ([700, 390, 129], [700], [390], [330], [320], [448], [512], [640]):
    s = synth(11, m_c, P=12, T=5)
    sire = np.array([0, 1, 2, 3, 4, 5, 6, 7, 3], np.int32)
    dam = np.array([1, 1, 5, 9, 10, 11, 6, 0, 8], np.int32)
    for model, add, dom in (("AD", s["add"], s["dom"]), ("AD", s["dom"], s["add"])):
        ref, _ = run(s, "dense", model, add, dom, sire, dam)
        for tol in (0.0, None):
            out, st = run(s, "int8_tensor", model, add, dom, sire, dam, tol=tol)
            d = np.abs(out - ref) / (np.abs(ref) + 1e-300)
            # slab [N][npairs][ncomp]: component 0 = VarA, 1 = VarD; pair 0 = (trait 1, trait 1)
            print(m_c, "swapped" if add is s["dom"] else "plain", "tol", tol, "fallback", st["last_fallback_units"],
                  "VarA max rel", float(d[:, :, 0].max()), "VarD max rel", float(d[:, :, 1].max()), flush=True)
