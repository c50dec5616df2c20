#!/usr/bin/env python
"""Executed warp instructions of one kernel grouped by named source blocks of encode.cu (markers = comment lines),
per block and per row of the C4 block.  usage: ncu_src_groups.py report.ncu-rep kernel_regex [rows]"""
import collections, csv, io, subprocess, sys
from pathlib import Path
rep, kern = sys.argv[1], sys.argv[2]
nrows = int(sys.argv[3]) if len(sys.argv) > 3 else 131072
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = ""; hdr = None; agg = collections.defaultdict(lambda: [0, 0])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] != "" and len(r) >= len(hdr):
        H = len(hdr); si = hdr.index("# Samples") - H; ii = hdr.index("Instructions Executed") - H
        try:
            a = agg[(cur, int(r[0]))]; a[0] += int(r[si] or 0); a[1] += int(r[ii] or 0)
        except ValueError: pass
ti = sum(a[1] for a in agg.values()); ts = sum(a[0] for a in agg.values())
src = (Path(__file__).resolve().parent.parent / "zdw_b200" / "csrc" / (sys.argv[4] if len(sys.argv) > 4 else "encode.cu")).read_text().split("\n")
marks = []
for i, l in enumerate(src):
    t = l.strip()
    if t.startswith("// ----") and len(t) > 12 and not t.startswith("// -----"): marks.append((i + 1, t[8:50]))
    elif ("__device__" in l or "__global__" in l) and "(" in l and not t.startswith("//"): marks.append((i + 1, "fn " + t.split("(")[0].split()[-1][:38]))
    elif t.startswith("k_") and "(" in t: marks.append((i + 1, "fn " + t.split("(")[0][:38]))
marks.sort()
tot = collections.defaultdict(lambda: [0, 0])
for (f, l), a in agg.items():
    if f != (sys.argv[4] if len(sys.argv) > 4 else "encode.cu"):
        k = "[" + f + "]"
    else:
        k = "encode.cu:top"
        for ln, name in marks:
            if l >= ln: k = f"{ln}: {name}"
            else: break
    tot[k][0] += a[0]; tot[k][1] += a[1]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k:52s} inst {100*v[1]/ti:5.1f}% ({v[1]/nrows:7.0f}/row) samples {100*v[0]/ts:5.1f}%")
print("total/row", round(ti / nrows), "launch instr", ti)
