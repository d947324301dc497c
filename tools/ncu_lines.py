#!/usr/bin/env python3
"""Rank the CUDA source lines of one kernel of an .ncu-rep by warp-stall samples.
usage: ncu_lines.py report.ncu-rep kernel-regex [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      "regex:" + kern, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = next(r for r in rows if "# Samples" in r)
i_s, i_ex = hdr.index("# Samples"), hdr.index("Instructions Executed")
cur, per = None, {}
for r in rows:
    if len(r) != len(hdr) or r is hdr:
        continue
    if r[0]:                                   # a CUDA source line; its SASS rows follow
        cur = (r[0], r[1].strip()[:120])
        per.setdefault(cur, [0, 0])
    elif cur is not None:
        try:
            per[cur][0] += int(r[i_s] or 0); per[cur][1] += int(r[i_ex] or 0)
        except ValueError:
            pass
tot = sum(v[0] for v in per.values()) or 1
print(f"{kern}: {tot} samples")
for (ln, src), (s, ex) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * s / tot:5.1f}%  L{ln:>4}  ex={ex:>6}  {src}")
