"""Small driver for ncu: one engine, cassava tiles, N crosses, one compute per mode (dense, int8_tensor, map_scan)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomicmateselectr_b200 import CrossVarEngine, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["int8_tensor", "map_scan"]
d = synth.make("cassava", N=N)
eng = CrossVarEngine(0)
eng.set_genmap(d["cM"], d["m_c"])
eng.set_haplotypes(d["H"])
ref = None
for mode in (["fp64"] + modes if "fp64" not in modes else modes):
    eng.set_mode(mode)
    eng.stage("AD", d["add"], d["dom"], None, d["sire"], d["dam"])
    eng.compute(sync=True)
    eng.compute(sync=True)
    if os.environ.get("PM_LOOP"):
        # sustained loop with SM clock / power sampling (nvidia-smi), to see throttling under this mode
        import subprocess, threading, time
        samples, stop = [], threading.Event()
        def sampler():
            while not stop.is_set():
                r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                                   capture_output=True, text=True).stdout.strip().split(",")
                try: samples.append((float(r[0]), float(r[1])))
                except Exception: pass
                time.sleep(0.05)
        th = threading.Thread(target=sampler); th.start()
        t0 = time.time(); ks = []
        while time.time() - t0 < float(os.environ["PM_LOOP"]):
            eng.compute(sync=True); ks.append(eng.stats()["last_ms_cross_kernel"])
        stop.set(); th.join()
        sm = sorted(x[0] for x in samples[2:]); pw = sorted(x[1] for x in samples[2:])
        if sm: print(f"  loop: kernel median {sorted(ks)[len(ks)//2]:.3f} ms over {len(ks)} runs; SM MHz median {sm[len(sm)//2]:.0f} min {sm[0]:.0f}; power W median {pw[len(pw)//2]:.0f} max {pw[-1]:.0f}")
    out, nseg = eng.fetch()
    st = eng.stats()
    if ref is None:
        ref = out
    rel = np.abs(out - ref) / np.maximum(np.abs(ref), 1e-300)
    err = float(rel.max())
    T = d["T"]
    var = np.abs(ref[:, :T, :])
    pairs = [(t, t) for t in range(T)] + [(a, b) for a in range(T) for b in range(a + 1, T)]
    scale = np.stack([np.sqrt(var[:, a, :] * var[:, b, :]) for a, b in pairs], axis=1)
    guarded = float((np.abs(out - ref) / np.maximum(np.abs(ref), 1e-6 * scale)).max())
    x, pr, c = np.unravel_index(np.argmax(rel), rel.shape)
    print(f"{mode:12s} cross kernel {st['last_ms_cross_kernel']:9.3f} ms  total {st['last_ms_total']:9.3f} ms  "
          f"max strict rel diff vs first mode {err:.2e} (pair {pr}, comp {c}: value {ref[x, pr, c]:.3e} vs scale {scale[x, pr, c]:.3e}; "
          f"abs diff / scale {abs(out[x, pr, c] - ref[x, pr, c]) / scale[x, pr, c]:.2e}); variances only {float(rel[:, :T, :].max()):.2e}; "
          f"guarded {guarded:.2e}; share > 1e-10: {float((rel > 1e-10).mean()):.2e}")
eng.close()
