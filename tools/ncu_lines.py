"""Developer helper: attribute the per-instruction samples of an ncu report to SOURCE LINES / functions.

Everything in the trajectory kernels is inlined from headers, and ncu's CSV source page only lists
SASS; this joins it (by instruction order) with `nvdisasm -g` of the same cubin, whose line-info
markers carry the inlining chain.

    cuobjdump -xelf all metrotrpl_b200/libmetrotrpl_b200.so        (in a scratch directory)
    python tools/ncu_lines.py report.ncu-rep scratch/team_kernels.sm_100a.cubin 'trpl_team_forward_kernelILi4ELi1ELb1'
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main(rep, cubin, kernel_pat, top=40, depth_inner=False, within=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    i_src, i_smp, i_exe = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    sass = [(r[i_src].strip(), int(r[i_smp] or 0), int(r[i_exe] or 0), [int(r[i] or 0) for i, _ in stall_cols])
            for r in rows[2:] if len(r) > i_exe]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and re.search(kernel_pat, l))
    # per instruction: the OUTERMOST frame below the kernel (the line of run_trajectory / finish_traj /
    # the kernel body the instruction was inlined from); `-gi` prints the inlining chain innermost first
    tags = []
    chain = []
    fresh = True
    for l in dis[start + 1:]:
        if l.startswith(".text.") or l.startswith(".section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            if fresh:
                chain = []
                fresh = False
            chain.append((m.group(1).split("/")[-1], int(m.group(2)), m.group(3).split("/")[-1] if m.group(3) else None))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            fresh = True
            outer = [c for c in chain if c[2] is not None and c[2].endswith(".cu")]
            if within is not None:
                # lines one frame below `within` (file:line of an outer frame): where that call spends its samples
                names = [f"{c[0]}:{c[1]}" for c in chain]
                if within in names:
                    k = names.index(within)
                    tags.append((chain[k - 1][0], chain[k - 1][1]) if k > 0 else (chain[0][0], chain[0][1]))
                else:
                    tags.append(("(elsewhere)", 0))
            elif depth_inner:
                tags.append((chain[0][0], chain[0][1]) if chain else ("?", 0))
            else:
                tags.append((outer[-1][0], outer[-1][1]) if outer else ((chain[-1][0], chain[-1][1]) if chain else ("?", 0)))
    n = min(len(tags), len(sass))
    if len(tags) != len(sass):
        print(f"warning: {len(tags)} disassembled instructions vs {len(sass)} profiled", file=sys.stderr)
    by_line = collections.Counter()
    by_file = collections.Counter()
    exe_line = collections.Counter()
    stall_line = collections.defaultdict(lambda: [0] * len(stall_cols))
    tot = 0
    for k in range(n):
        f, ln = tags[k]
        s = sass[k][1]
        by_line[(f, ln)] += s
        by_file[f] += s
        exe_line[(f, ln)] += sass[k][2]
        for j, v in enumerate(sass[k][3]):
            stall_line[(f, ln)][j] += v
        tot += s
    print(f"total samples {tot}")
    for f, s in by_file.most_common():
        print(f"  {f:24s} {100.0 * s / tot:5.1f}%")
    print("top lines:")
    for (f, ln), s in by_line.most_common(top):
        st = stall_line[(f, ln)]
        best = sorted(zip(st, [h for _, h in stall_cols]), reverse=True)[:3]
        print(f"  {f}:{ln:<5d} {100.0 * s / tot:5.1f}%  exe {exe_line[(f, ln)]:>11d}  " +
              " ".join(f"{h[6:]}={v}" for v, h in best if v))
    return by_line, tot


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40, "--inner" in sys.argv,
         next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--within=")), None))
