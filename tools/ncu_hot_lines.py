#!/usr/bin/env python
"""Per-source-line executed-instruction and stall shares of one launch of an .ncu-rep captured with --import-source on
(kernels built with -lineinfo).  usage: ncu_hot_lines.py report.ncu-rep launch_index [top_n]"""
import collections, csv, io, subprocess, sys

rep, li = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--launch-skip", li,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur, H, name = None, None, None
agg, stall, text = collections.Counter(), collections.Counter(), {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) == 2 and r[0] == "Function Name":
        name = r[1]
        continue
    if r and r[0] == "Line No":
        H = r
        continue
    if H is None or len(r) < len(H):
        continue
    try:
        ln, n = int(r[0]), int(r[H.index("Instructions Executed")])
    except ValueError:
        continue
    agg[(cur, ln)] += n
    text[(cur, ln)] = r[1]
    try:
        stall[(cur, ln)] += int(r[H.index("Warp Stall Sampling (All Samples)")])
    except ValueError:
        pass
tot, ts = sum(agg.values()), max(sum(stall.values()), 1)
print(name)
print(f"{'file:line':28s} {'instr%':>7s} {'stall%':>7s}  source")
for (f, ln), n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{f + ':' + str(ln):28s} {n / tot * 100:6.1f}% {stall[(f, ln)] / ts * 100:6.1f}%  {text[(f, ln)].strip()[:110]}")
