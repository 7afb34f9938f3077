#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per kernel the metrics DESIGN.md / profiles/ quote.
usage: python tools_ncu_summary.py gpurun_out/x.ncu-rep [--top-sass KERNEL_REGEX]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible"),
    ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "st_long_sb"),
    ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "st_short_sb"),
    ("smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "st_mio"),
    ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "st_lg"),
    ("smsp__warp_issue_stalled_wait_per_warp_active.pct", "st_wait"),
    ("smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "st_branch"),
    ("smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "st_notsel"),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "st_barrier"),
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].split("::")[-1]
        print("== %s" % name)
        for m, short in WANT:
            if m in idx:
                print("   %-16s %s %s" % (short, r[idx[m]], units[idx[m]]))


def sass(rep, regex, top=45):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex],
                         capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(out))]
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    ie, av, src, smp = idx["Instructions Executed"], idx["Avg. Threads Executed"], idx["Source"], idx["# Samples"]
    body = [r for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
    tot = sum(int(r[ie]) for r in body)
    tots = sum(int(r[smp]) for r in body)
    print("total warp instructions %d, samples %d" % (tot, tots))
    for r in sorted(body, key=lambda r: -int(r[smp]))[:top]:
        print("%6.2f%% samp  %10d inst  thr=%4s  %s" % (100.0 * int(r[smp]) / max(tots, 1), int(r[ie]), r[av], r[src][:90]))


if __name__ == "__main__":
    if "--top-sass" in sys.argv:
        sass(sys.argv[1], sys.argv[sys.argv.index("--top-sass") + 1])
    else:
        raw(sys.argv[1])
