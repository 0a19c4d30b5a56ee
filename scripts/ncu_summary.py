#!/usr/bin/env python
"""Condense an ncu report (--set full) and a launch list (gpu__time_duration.sum CSV) into a small
text summary for profiles/.  Usage: ncu_summary.py <prof.ncu-rep> <launches.csv> <out.md> [title]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "smsp__dyn_smem_size_per_block", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def main():
    rep, launches, out = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else rep
    lines = [f"# {title}", ""]
    if rep != "-":
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            lines += [f"## ncu --set full: `{name[:100]}`", "", "| metric | unit | value |", "|---|---|---|"]
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    lines.append(f"| {k} | {units[i]} | {r[i]} |")
            lines.append("")
    if launches != "-":
        per = defaultdict(list)
        for r in csv.DictReader(l for l in open(launches) if l.startswith('"')):
            if r.get("Metric Name") == "gpu__time_duration.sum":
                per[r["Kernel Name"].split("(")[0][:90]].append(float(r["Metric Value"]))
        tot = sum(sum(v) for v in per.values())
        lines += ["## launch list (ncu --metrics gpu__time_duration.sum; cold-cache, serialised: compare shares)", "",
                  "| kernel | launches | total ms | mean ms | share |", "|---|---|---|---|---|"]
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            lines.append(f"| {k} | {len(v)} | {sum(v) / 1e6:.3f} | {sum(v) / len(v) / 1e6:.4f} | {sum(v) / tot:.1%} |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
