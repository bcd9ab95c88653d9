#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small markdown file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_sgp4.ncu-rep profiles/r01_sgp4_grid.md [cells_per_launch]
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    cells = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    lines = [f"# ncu summary of `{rep.split('/')[-1]}`", "",
             "Captured with `ncu --set full --clock-control none --import-source on` under gpurun (one B200); read on the "
             "CPU box with `ncu -i ... --page raw|source --csv`. Per-launch values, one column per captured launch.", ""]
    kcol = hdr.index("Kernel Name")
    lines.append("Kernels: " + "; ".join(sorted({r[kcol] for r in rows})))
    lines += ["", "| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows))) + " |",
              "|---|---|" + "---|" * len(rows)]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"| `{k}` | {units[i]} | " + " | ".join(r[i] for r in rows) + " |")
    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    h = None
    ops, samples, total = collections.Counter(), collections.Counter(), 0
    for r in src:
        if "Source" in r and "Instructions Executed" in r:
            h = r
            continue
        if h is None or len(r) < len(h):
            continue
        n = r[h.index("Instructions Executed")]
        if not n.isdigit():
            continue
        m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[h.index("Source")].strip())
        op = m.group(2).split(".")[0] if m else "?"
        ops[op] += int(n)
        samples[op] += int(r[h.index("Warp Stall Sampling (All Samples)")] or 0)
        total += int(n)
    if total:
        nl = max(len(rows), 1)
        # the source page accumulates over captured launches and replay passes: normalise its totals to the
        # hardware counter smsp__inst_executed.sum of one launch
        if "smsp__inst_executed.sum" in hdr:
            one = float(rows[0][hdr.index("smsp__inst_executed.sum")].replace(",", ""))
            nl = total / one
        lines += ["", f"## Executed warp-instructions by opcode (source page, normalised to one launch via "
                      "smsp__inst_executed.sum)", ""]
        per = (lambda n: n / nl / (cells / 32.0)) if cells else None
        lines.append("| opcode | warp-instr / launch" + (" | per cell |" if cells else " |") + " stall samples |")
        lines.append("|---|---|" + ("---|" if cells else "") + "---|")
        for op, n in ops.most_common(24):
            lines.append(f"| {op} | {n / nl:.0f} | " + (f"{per(n):.1f} | " if cells else "") + f"{samples[op]} |")
        lines.append(f"| **total** | {total / nl:.0f} | " + (f"{per(total):.1f} | " if cells else "") + f"{sum(samples.values())} |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
