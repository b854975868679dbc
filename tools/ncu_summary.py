#!/usr/bin/env python
"""Summarises an Nsight Compute report (read here, no GPU needed) into the text files kept under
profiles/: the headline raw metrics per profiled launch and the hottest source lines.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_fused.txt [--top 40]
    python tools/ncu_summary.py --launches gpurun_out/launches.csv profiles/r01_launches.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

csv.field_size_limit(10 ** 9)

RAW = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__shared_mem_per_block_allocated",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def summarize_report(rep, out, top):
    lines = ["# ncu summary of %s" % rep, ""]
    rows = list(csv.reader(ncu(["-i", rep, "--page", "raw", "--csv"]).splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        lines.append("## launch %s: %s  grid %s block %s" % (r[col["ID"]], r[col["Kernel Name"]][:100],
                                                           r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]))
        for m in RAW:
            if m in col:
                lines.append("  %-82s %14s %s" % (m, r[col[m]], units[col[m]]))
        lines.append("")
    src = list(csv.reader(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]).splitlines()))
    cur, agg = None, defaultdict(lambda: [0, 0, ""])
    kernel_seen = 0
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) >= 2 and r[0] == "Kernel Name":
            kernel_seen += 1
        if len(r) < 8 or r[0] in ("Line No", "Function Name", ""):
            continue
        try:
            ln, smp, ins = int(r[0]), int(r[4]), int(r[7])
        except ValueError:
            continue
        a = agg[(cur, ln)]
        a[0] += ins
        a[1] += smp
        a[2] = r[1].strip()[:100]
    ti = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    lines.append("## hottest source lines (all profiled launches summed): %% of warp instructions executed, "
                 "%% of stall samples")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        lines.append("  %5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * a[0] / ti, 100.0 * a[1] / ts, f, ln, a[2]))
    lines.append("")
    lines.append("## most stalled source lines")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top // 2]:
        lines.append("  %5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * a[1] / ts, 100.0 * a[0] / ti, f, ln, a[2]))
    open(out, "w").write("\n".join(lines) + "\n")


def summarize_launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    grid, block = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = defaultdict(lambda: [0, 0.0, "", ""])
    total = 0.0
    for r in rows[1:]:
        try:
            t = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        name = r[kn].split("(")[0].replace("void ", "")
        a = agg[name]
        a[0] += 1
        a[1] += t
        a[2], a[3] = r[grid], r[block]
        total += t
    lines = ["# ncu launch list %s: per-kernel device time (gpu__time_duration.sum, ns; cold-cache, serialised)" % path,
             "# kernel | launches | total ns | mean ns | share | last grid | last block"]
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-60s %5d %12.0f %10.0f %6.1f%%  %s %s" % (name[:60], a[0], a[1], a[1] / a[0],
                                                                 100 * a[1] / total, a[2], a[3]))
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    a = sys.argv[1:]
    top = 40
    if "--top" in a:
        i = a.index("--top")
        top = int(a[i + 1])
        del a[i:i + 2]
    if a[0] == "--launches":
        summarize_launches(a[1], a[2])
    else:
        summarize_report(a[0], a[1], top)
