#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep
(needs -lineinfo and --import-source on):  python scripts/ncu_lines.py x.ncu-rep [top]"""
import csv, subprocess, sys
from collections import defaultdict
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None
agg = defaultdict(lambda: [0, 0, ""])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if r[0] != "":  # a source line summary row
        key = (cur_file, int(r[0]))
        agg[key][0] += int(r[ie]) if r[ie].isdigit() else 0; agg[key][1] += int(r[isamp]) if r[isamp].isdigit() else 0; agg[key][2] = r[1].strip()
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total warp instructions {tot}, samples {ts}")
for (f, ln), (n, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n/tot*100:6.2f}% inst {s/ts*100:6.2f}% smp  {f}:{ln:<4d} {src[:110]}")
