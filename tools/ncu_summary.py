"""Text summary of one kernel of an .ncu-rep for profiles/ (read here, on the CPU container; the report itself stays in gpurun_out/):
    python tools/ncu_summary.py gpurun_out/<name>.ncu-rep > profiles/<name>.txt
Prints the counters DESIGN.md quotes (raw page) and the source page aggregated by SASS opcode (warp stall samples and
executed warp instructions; BRA = mbarrier try_wait spin loops)."""
import collections
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__cluster_dim_x", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_branch_resolving"]


def page(rep, name):
    return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], check=True, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(page(rep, "raw"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:100s} {vals[i]:>20s} {units[i]}")
    src = list(csv.reader(io.StringIO(page(rep, "source"))))
    h = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    cols = src[h]
    i_src, i_smp, i_exe = cols.index("Source"), cols.index("# Samples"), cols.index("Instructions Executed")
    agg = collections.defaultdict(lambda: [0, 0])
    total = n = 0
    for r in src[h + 1:]:
        if len(r) <= i_exe:
            continue
        toks = r[i_src].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0]
        smp, exe = int(r[i_smp] or 0), int(r[i_exe] or 0)
        agg[op][0] += smp
        agg[op][1] += exe
        total += smp
        n += 1
    print(f"\n# source page aggregated by opcode: total samples {total}, {n} SASS instructions")
    for op, (smp, exe) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
        print(f"{op:12s} samples {smp:9d} {100.0 * smp / max(total, 1):5.1f}%  executed {exe}")


if __name__ == "__main__":
    main()
