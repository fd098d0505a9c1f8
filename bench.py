#!/usr/bin/env python
"""Benchmark of the cross-variance hot path (BASELINE.json metric):
cross variance-covariance predictions/sec (A+AD, 5 traits) on synthetic cassava-scale data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cassava]

A step = one pass of the hot path (modelType AD: VarA + VarD for all 15 trait pairs + Nsegsnps) over
the whole batch of crosses.  N > 1: launched by torchrun, one rank per GPU, crosses sharded by rank
(weak scaling: --crosses-per-gpu fixed), LD-decay tiles and haplotypes replicated, no data-path
collective; the per-rank slabs are gathered to rank 0 in the e2e leg only.

  value     crosses/s with inputs resident in HBM (gms_compute only), CUDA events, max over ranks
  e2e       the same through gms_predict with pinned HOST buffers (H2D + kernels + D2H [+ gather])
  roofline  FP64 tensor (DMMA) roofline of the dominant kernel (per-cross dominance contraction)
  cpu_baseline  the oracle's C restatement of the reference algorithm on the host cores (rank 0)
--impl reference times that CPU restatement alone (the reference is R; R is not installed here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cross variance-covariance predictions/sec (A+AD, 5 traits)"
UNIT = "crosses/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cassava", choices=["cassava", "large", "mini"])
    ap.add_argument("--crosses-per-gpu", type=int, default=None, help="default: the workload's N on every GPU")
    ap.add_argument("--model", default="AD", choices=["A", "AD", "DirDom"])
    ap.add_argument("--cpu-sample", type=int, default=8, help="crosses per CPU-baseline step (3 steps, ~20 s on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        t0, t1 = getattr(self, "t0", None), getattr(self, "t1", None)
        inside = [ln for (ts, ln) in self.lines if t0 is None or (t0 <= ts <= t1 + 0.06)]
        if not inside:          # timed region shorter than one sampling period: the closest samples under load
            inside = [ln for (ts, ln) in self.lines if t1 is None or ts <= t1 + 0.06][-3:]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def dgemm_peak_tflops(torch):
    """FP64 ceiling on this box: cuBLAS DGEMM 8192^3, best of 5 (MEASURED_PEAKS.json has no FP64 figure)."""
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * 8192 ** 3 / best / 1e9


def int8_peak_tops(torch):
    """int8 tensor ceiling on this box: cuBLASLt int8 GEMM 8192^3 through torch._int_mm, best of 5."""
    try:
        a = torch.randint(-100, 100, (8192, 8192), dtype=torch.int8, device="cuda")
        b = torch.randint(-100, 100, (8192, 8192), dtype=torch.int8, device="cuda").t().contiguous().t()
        for _ in range(2):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * 8192 ** 3 / best / 1e9
    except Exception:
        return 4500.0      # nominal dense int8, if the library call is unavailable


# ----------------------------------------------------------------------------------------------
def cpu_reference_run(d, model, sample, steps, warmup, nthreads=0):
    """Time the oracle's C restatement of the reference algorithm (dense D per cross, one quadratic
    form per trait pair) on the host cores: `sample` crosses per step, all threads inside a cross."""
    from oracle import crossvar_ref as CR
    from oracle import predcrossvar_oracle as O
    offs = np.concatenate([[0], np.cumsum(d["m_c"])])
    blocks = []
    for c in range(len(d["m_c"])):
        cm = d["cM"][offs[c]:offs[c + 1]]
        blocks.append(O.recombfreq_to_ld_decay(O.genmap2recombfreq(cm, np.ones(cm.size))))
    prob = CR.RefProblem(d["H"], d["add"], d["dom"] if model != "A" else None, blocks=blocks)
    # all host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    cores = nthreads or max(avail, CR.max_threads())
    rng = np.random.default_rng(12345)
    times, nsegs = [], []
    for it in range(warmup + steps):
        pick = rng.integers(0, d["N"], sample)
        t0 = time.perf_counter()
        for x in pick:
            _, nseg = prob.one_cross(int(d["sire"][x]), int(d["dam"][x]), "AD" if model != "A" else "A", cores)
            nsegs.append(nseg)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = float(np.sum(times))
    return {"value": sample * steps / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample} random crosses per step x {steps} steps of the same workload "
                      f"(mean Nsegsnps {int(np.mean(nsegs))}); oracle/crossvar_ref.c = the reference's algorithm "
                      f"(dense m_seg^2 cross LD matrix, one quadratic form per trait pair) restated in C+OpenMP; "
                      f"R itself is not installed",
            "ms_per_step": 1e3 * total / steps}


_REAL_STDOUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line: route everything else (NCCL's version banner, library
    chatter) to stderr and keep the real stdout for emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    args = parse()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from genomicmateselectr_b200 import synth
    cfg = synth.CONFIGS[args.workload]
    n_per = args.crosses_per_gpu or cfg["N"]
    T = cfg["T"]
    config = {"workload": f"synthetic {args.workload}-scale: {cfg['P']} phased parents, {sum(cfg['m_c'])} SNPs / "
                          f"{len(cfg['m_c'])} chromosomes, {T} traits, {n_per} crosses per GPU, modelType {args.model}",
              "model_type": args.model, "crosses_per_gpu": n_per, "parents": cfg["P"], "snps": int(sum(cfg["m_c"])),
              "chromosomes": len(cfg["m_c"]), "traits": T, "sharding": "by cross, tiles+haplotypes replicated",
              "l2_policy": "inputs larger than L2 (RoR tile image 283 MB at cassava scale, re-streamed every step)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        d = synth.make(args.workload, N=min(cfg["N"], 50_000))
        r = cpu_reference_run(d, args.model, args.cpu_sample, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist
    from genomicmateselectr_b200 import CrossVarEngine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_total = n_per * world
    d = synth.make(args.workload, N=n_total)
    lo, hi = rank * n_per, (rank + 1) * n_per
    sire, dam = d["sire"][lo:hi].copy(), d["dam"][lo:hi].copy()
    dom = d["dom"] if args.model != "A" else None
    asub = (d["add"] + d["dom"] * 0.1) if args.model == "DirDom" else None

    eng = CrossVarEngine(local_rank)
    stream = torch.cuda.Stream()          # a real (non-default) stream: handle 0 would mean "engine's own stream"
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    eng.set_genmap(d["cM"], d["m_c"])
    eng.set_haplotypes(d["H"])
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    fp64_peak = dgemm_peak_tflops(torch)
    int8_peak = int8_peak_tops(torch)

    # pinned host buffers for the e2e legs
    npairs, ncomp = T * (T + 1) // 2, {"A": 1, "AD": 2, "DirDom": 3}[args.model]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    add_p = pin(np.asfortranarray(d["add"]).T)           # m x T column-major == (T, m) C-order
    dom_p = pin(np.asfortranarray(dom).T) if dom is not None else None
    asub_p = pin(np.asfortranarray(asub).T) if asub is not None else None
    sire_p, dam_p = pin(sire), pin(dam)
    out_p = torch.empty((n_per, npairs, ncomp), dtype=torch.float64).pin_memory()
    nseg_p = torch.empty(n_per, dtype=torch.int32).pin_memory()
    gather = [torch.empty((n_per, npairs, ncomp), dtype=torch.float64, device="cuda") for _ in range(world)] \
        if (world > 1 and rank == 0) else None
    f = lambda t: None if t is None else t.numpy().T      # back to an F-ordered (m, T) view of the pinned memory

    def e2e_step():
        o, n = eng.predict(args.model, f(add_p), f(dom_p), f(asub_p), sire_p.numpy(), dam_p.numpy(),
                           out=out_p.numpy(), nseg=nseg_p.numpy())
        if world > 1:        # the "final gather of results" of the north star
            dist.gather(out_p.cuda(non_blocking=True), gather, dst=0)
        return float(o[0, 0, 0])

    def measure(mode, with_clocks=False):
        """One evaluation mode on the same staged inputs: device-resident throughput (gms_compute only: every
        kernel incl. the device-side operand preparation), per-kernel times, and end-to-end through gms_predict."""
        eng.set_mode(mode)
        eng.stage(args.model, d["add"], dom, asub, sire, dam)
        for _ in range(args.warmup):
            eng.compute(sync=True)
        sampler = ClockSampler(local_rank) if with_clocks else None
        if sampler:
            sampler.start()       # nvidia-smi needs ~0.3 s to start: run it through the warm-up, keep the timed window
            tw = time.time()
            while len(sampler.lines) < 2 and time.time() - tw < 3.0:
                eng.compute(sync=True)
        barrier()
        w0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            eng.compute(sync=False)
        e1.record(stream)
        barrier()
        w1 = time.time()
        ms = max_over_ranks(e0.elapsed_time(e1))
        clocks = None
        if sampler:
            sampler.window(w0, w1)
            clocks = sampler.stop()
        # dominant-kernel duration: CUDA events recorded around the launch on the same stream (inside the
        # library); averaged over separate, synchronised steps so that every launch is read back
        kernel_ms, stage_ms = [], []
        for _ in range(args.steps):
            eng.compute(sync=True)
            s_ = eng.stats()
            kernel_ms.append(s_["last_ms_cross_kernel"])
            stage_ms.append((s_["last_ms_parent_stage"], s_["last_ms_cross_stage"], s_["last_ms_assemble"], s_["last_ms_total"]))
        st_ = eng.stats()
        for _ in range(min(2, args.warmup)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        return {"mode": mode, "resolved_mode": {0: "fp64_dmma", 1: "map_scan", 2: "int8_tensor"}[st_["last_mode"]],
                "value": n_total * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps,
                "kernel_ms": float(np.mean(kernel_ms)), "step_ms_sync": float(np.mean([s_[3] for s_ in stage_ms])),
                "stages_ms": {"prep_and_parent_tables": float(np.mean([s_[0] for s_ in stage_ms])),
                              "cross_term": float(np.mean([s_[1] for s_ in stage_ms])),
                              "assemble_nseg": float(np.mean([s_[2] for s_ in stage_ms]))},
                "launches": st_["last_launches"] * args.steps, "e2e_value": n_total * args.steps / e2e_s,
                "e2e_ms_per_step": 1e3 * e2e_s / args.steps, "clocks": clocks, "stats": st_}

    def fp64_roofline(r):
        """FP64 tensor (DMMA) roofline of contract_kernel<MASK_XI>."""
        if args.model == "A" or r["kernel_ms"] <= 0:
            return None
        alg = 2.0 * T * r["stats"]["S2"] * n_per            # SURVEY 8d dense count for the cross term, per launch
        exe = r["stats"]["last_flops_cross"]
        ach = alg / (r["kernel_ms"] * 1e-3) / 1e12
        return {"bound": "tensor", "kernel": "contract_kernel<MASK_XI> (FP64 mma.sync DMMA, cp.async.bulk staged)",
                "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": traffic,
                "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 figure); "
                               "DMMA pipe probe tools/dmma_probe: 37.2 TFLOP/s; nominal 37-40",
                "algorithmic_flops_per_launch": alg, "executed_flops_per_launch": exe,
                "executed_tflops": exe / (r["kernel_ms"] * 1e-3) / 1e12,
                "executed_frac_of_peak": exe / (r["kernel_ms"] * 1e-3) / 1e12 / fp64_peak,
                "note": "algorithmic = dense 2*T*S2 per cross (no symmetry credit); the kernel executes only the lower "
                        "block triangle of the symmetric RoR blocks, so frac can exceed 1",
                "kernel_ms": r["kernel_ms"], "kernel_share_of_step": r["kernel_ms"] / r["step_ms_sync"]}

    def int8_roofline(r):
        """int8 tensor (tcgen05) roofline of i8_contract_kernel: 7 exact digit planes x ternary xi."""
        if args.model == "A" or r["kernel_ms"] <= 0:
            return None
        mc = np.asarray(d["m_c"], dtype=np.float64)
        nJ = np.ceil(mc / 64.0)
        alg = 2.0 * 7 * T * n_per * float(np.sum(mc * (mc + 1) / 2))                    # lower triangle incl. diagonal, 7 digits
        exe = 2.0 * 7 * T * (np.ceil(n_per / 128.0) * 128) * float(np.sum(64 * 64 * nJ * (nJ + 1) / 2))
        ach = alg / (r["kernel_ms"] * 1e-3) / 1e12
        return {"bound": "tensor", "kernel": "i8_contract_kernel (tcgen05.mma kind::i8, int32 accumulators in TMEM, cp.async.bulk staged)",
                "achieved": ach, "peak": int8_peak, "unit": "TFLOP/s", "op_type": "int8 multiply-add = 2 ops",
                "frac": ach / int8_peak, "traffic": traffic_i8,
                "peak_source": "cuBLASLt int8 GEMM 8192^3 (torch._int_mm) measured in this run (MEASURED_PEAKS.json has bf16 only); "
                               "nominal dense int8 4500",
                "frac_of_nominal_int8": ach / 4500.0,
                "algorithmic_flops_per_launch": alg, "executed_flops_per_launch": exe,
                "executed_tflops": exe / (r["kernel_ms"] * 1e-3) / 1e12,
                "executed_frac_of_peak": exe / (r["kernel_ms"] * 1e-3) / 1e12 / int8_peak,
                "fp64_dense_equivalent_tflops": 2.0 * T * r["stats"]["S2"] * n_per / (r["kernel_ms"] * 1e-3) / 1e12,
                "note": "algorithmic = 7 radix-128 digit planes x the exact lower triangle (incl. diagonal) of every chromosome "
                        "block x T traits x crosses; executed adds 64-granular block and 128-cross tile padding",
                "kernel_ms": r["kernel_ms"], "kernel_share_of_step": r["kernel_ms"] / r["step_ms_sync"]}

    # DRAM bytes per launch of the two contraction kernels, from the committed `ncu --set full` captures of this config
    traffic = traffic_i8 = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if (tr["workload"], tr["crosses_per_gpu"], tr["model"]) == (args.workload, n_per, args.model):
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            if "int8" in tr:
                traffic_i8 = tr["int8"]["dram_bytes_read"] + tr["int8"]["dram_bytes_write"]
    except Exception:
        pass

    head = measure("auto", with_clocks=True)                 # the default path of the library = the headline
    ms, value, launches, clocks, st = head["ms_per_step"] * args.steps, head["value"], head["launches"], head["clocks"], head["stats"]
    roof = int8_roofline(head) if head["resolved_mode"] == "int8_tensor" else fp64_roofline(head)
    h2d = sum(t.numel() * t.element_size() for t in (add_p, dom_p, asub_p, sire_p, dam_p) if t is not None) \
        + 2 * n_per * 4 + 4 * int(st["last_parents_used"])      # + sire/dam table slots + parent list
    d2h = out_p.numel() * 8 + nseg_p.numel() * 4
    e2e = {"value": head["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": head["e2e_ms_per_step"],
           "api": "gms_predict (C ABI) via CrossVarEngine.predict, pinned host buffers, default mode"}

    def leg(mode, note, roofline_fn=None):
        try:
            r = measure(mode)
            out = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "cross_kernel_ms": r["kernel_ms"],
                   "e2e": r["e2e_value"], "stages_ms": r["stages_ms"], "note": note}
            if roofline_fn:
                out["roofline"] = roofline_fn(r)
            return out
        except Exception as exc:   # the headline path must not depend on the other legs
            return {"error": str(exc)}

    fp64_leg = leg("fp64", "gms_set_mode(GMS_MODE_DENSE): the FP64 DMMA contraction north_star describes; any symmetric matrix",
                   fp64_roofline) if head["resolved_mode"] != "fp64_dmma" else None
    map_scan = leg("map_scan", "gms_set_mode(GMS_MODE_MAP_SCAN): map-structured prefix-sum path, map-derived recombFreqMat only; opt-in")
    eng.set_mode("auto")

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference_run(d, args.model, args.cpu_sample, 3, 1)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        pk = measured_peaks()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None,
                "dtype": "f64 results; " + ("exact int8 digit planes on tcgen05 + f64 recombination" if head["resolved_mode"] == "int8_tensor" else "f64 DMMA"),
                "data": "synthetic", "config": dict(config, mode=head["resolved_mode"]), "clocks": clocks,
                "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
                "stages_ms": head["stages_ms"], "fp64_dmma": fp64_leg, "map_scan": map_scan,
                "setup_s": setup_s, "hbm_tile_bytes": int(st["tile_bytes"]),
                "measured_peaks": {"hbm_gbs": pk.get("hbm_gbs"), "fp64_dgemm_tflops": fp64_peak, "int8_gemm_tops": int8_peak}}
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
