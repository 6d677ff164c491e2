#!/usr/bin/env python
"""Condense `ncu --page raw --csv` / `--page source --csv` exports (made on the GPU box by tools/gpu_round.sh)
into the small tracked summaries under profiles/.

    python tools/ncu_summary.py gpurun_out/<tag>_prof_fused profiles/<tag>_ncu_fused.md
"""
import collections
import csv
import re
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "memory_l1_wavefronts_shared_ideal",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
]


def main(base, out):
    lines = [f"# ncu summary of `{base}` (--set full --clock-control none, exported on the GPU box)\n"]
    rows = list(csv.reader(open(base + "_raw.csv")))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        lines.append(f"\n## {r[col['Kernel Name']]}\n")
        lines.append("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in col:
                lines.append(f"| {k} | {r[col[k]]} | {units[col[k]]} |")
        tr = float(r[col["dram__bytes_read.sum"]]) + float(r[col["dram__bytes_write.sum"]])
        lines.append(f"| dram traffic (read+write) | {tr:.4f} | {units[col['dram__bytes_read.sum']]} |")
        lines.append("\nstall reasons (warps stalled per issue-active cycle, > 0.15):\n")
        for h, i in col.items():
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio", h)
            if m:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.15:
                    lines.append(f"* {m.group(1)}: {v:.2f}")
    try:
        rows = list(csv.reader(open(base + "_source.csv")))
    except FileNotFoundError:
        rows = []
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    seen = set()
    for k, s in enumerate(starts):
        e = starts[k + 1] if k + 1 < len(starts) else len(rows)
        sec = rows[s + 1:e]
        if not sec or len(sec[0]) < 2 or sec[0][1] != "Source":
            continue
        name = rows[s][1]
        if name in seen:
            continue
        h = sec[0]
        iE, iS = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
        ops, stalls, tot = collections.Counter(), collections.Counter(), 0
        for r in sec[1:]:
            if len(r) <= iE or not r[0].startswith("0x"):
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1])
            op = m.group(2).split(".")[0] if m else "?"
            n = int(r[iE])
            ops[op] += n
            tot += n
            stalls[op] += int(r[iS])
        if not tot:
            continue
        seen.add(name)
        ts = max(1, sum(stalls.values()))
        lines.append(f"\n## SASS opcode mix of {name}\n\n{tot} warp instructions\n")
        lines.append("| opcode | warp-instructions | share | stall samples |\n|---|---|---|---|")
        for op, n in ops.most_common(24):
            lines.append(f"| {op} | {n} | {100 * n / tot:.1f}% | {100 * stalls[op] / ts:.1f}% |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
