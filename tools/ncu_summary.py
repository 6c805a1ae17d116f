#!/usr/bin/env python
"""Summarise one-launch `ncu --set full` reports into the text files kept under profiles/.

usage: python tools/ncu_summary.py REPORT.ncu-rep "header line" [more header lines...] > profiles/rNN_ncu_full_X.txt
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "launch__grid_size",
    "launch__block_size",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__cycles_elapsed.avg",
    "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def raw_page(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True)
    rows = list(csv.reader(io.StringIO(out.stdout)))
    names, units, vals = rows[0], rows[1], rows[2]
    return {n: (v, u) for n, u, v in zip(names, units, vals)}


def sass_mix(report, top=14):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True)
    if out.returncode:
        return []
    rows = list(csv.reader(io.StringIO(out.stdout)))
    if not rows:
        return []
    hi = next((i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r), None)
    if hi is None:
        return []
    h = rows[hi]
    si = h.index("Source")
    ei = h.index("Instructions Executed")
    mix = {}
    for r in rows[hi + 1:]:
        if len(r) <= max(si, ei):
            continue
        toks = r[si].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.rstrip(";")
        try:
            mix[op] = mix.get(op, 0) + int(float(r[ei].replace(",", "")))
        except ValueError:
            pass
    tot = sum(mix.values()) or 1
    return [(k, v, 100.0 * v / tot) for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:top]]


def main():
    report = sys.argv[1]
    m = raw_page(report)
    kname = m.get("Kernel Name", ("?", ""))[0]
    print(f"# ncu --set full --clock-control none --import-source on, one launch of {kname}")
    for line in sys.argv[2:]:
        print(f"# {line}")
    print()
    for k in KEEP:
        if k in m:
            v, u = m[k]
            print(f"{k:100s} {v:>16s} {u}")
    mix = sass_mix(report)
    if mix:
        print("\n# warp-level SASS mix (source page, 'Instructions Executed'), top opcodes")
        for op, n, pct in mix:
            print(f"{op:28s} {n:>14d} {pct:6.2f} %")


if __name__ == "__main__":
    main()
