"""Summarise an .ncu-rep (raw page + per-instruction source page) into a small markdown file.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.md "title"
"""
import collections
import csv
import io
import re
import subprocess
import sys


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, dst, title):
    raw = ncu_csv(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
            "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_local_ld.sum",
            "smsp__sass_inst_executed_op_local_st.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    lines = [f"# {title}", "", f"source report: `{rep}` (ncu --set full --clock-control none --import-source on)", "",
             "| metric | value | unit |", "|---|---|---|"]
    for k in keys:
        if k in d:
            lines.append(f"| {k} | {d[k][0]} | {d[k][1]} |")
    lines += ["", "## warp stall reasons (cycles per issued instruction)", "", "| reason | value |", "|---|---|"]
    for k in hdr:
        if "issue_stalled" in k and "per_issue_active" in k:
            lines.append(f"| {k.split('stalled_')[1].split('_per')[0]} | {float(d[k][0]):.3f} |")
    src = ncu_csv(rep, "source", ("--print-source", "sass"))
    sh = src[1]
    ix = {h: i for i, h in enumerate(sh)}
    byop = collections.Counter()
    tot = 0
    for r in src[2:]:
        if len(r) != len(sh):
            continue
        m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]].strip())
        n = int(r[ix["Instructions Executed"]])
        byop[m.group(1) if m else "?"] += n
        tot += n
    lines += ["", "## executed warp instructions by opcode", "", f"total {tot}", "", "| opcode | share |", "|---|---|"]
    for op, n in byop.most_common(20):
        lines.append(f"| {op} | {100 * n / tot:.1f}% |")
    fp64 = sum(byop[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
    lines.append(f"\nFP64-pipe share of executed instructions: {100 * fp64 / tot:.1f}%")
    open(dst, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:4])
