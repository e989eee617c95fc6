"""Per-source-line hot spots of one kernel launch from an .ncu-rep (compiled with -lineinfo, captured with --import-source on).
usage: python profiles/ncu_lines.py <report> <kernel regex> <launch-skip> [top]"""
import csv
import subprocess
import sys

rep, kern, skip = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, lines = "", None, []
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":
        lines.append((fname, r))
i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
stall = {n: hdr.index(n) for n in ("stall_wait", "stall_short_sb", "stall_long_sb", "stall_barrier", "stall_branch_resolving", "stall_mio", "stall_lg")}
tot_s = sum(int(r[i_s] or 0) for _, r in lines) or 1
tot_i = sum(int(r[i_i] or 0) for _, r in lines) or 1
print(f"total warp instructions {tot_i:,}  samples {tot_s:,}")
for f, r in sorted(lines, key=lambda x: -int(x[1][i_s] or 0))[:top]:
    ins = int(r[i_i] or 0)
    st = sorted(((int(r[j] or 0), n) for n, j in stall.items()), reverse=True)[:2]
    print(f"{int(r[i_s]) * 100 / tot_s:5.1f}% samp {ins * 100 / tot_i:5.1f}% inst lanes {int(r[i_t] or 0) / max(1, ins):4.1f} {st[0][1][6:]:>10} | {f[:12]}:{r[0]:>4} {r[1].strip()[:120]}")
