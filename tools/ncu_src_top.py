#!/usr/bin/env python
"""Top CUDA source lines by warp-stall samples of one kernel in an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_src_top.py report.ncu-rep kernel_regex [N]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = ""
hdr = None
lines = []
seen_fn = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        seen_fn += 1; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and r[0] != "":
        lines.append((cur_file, r))
if not hdr:
    sys.exit("no source table")
H = len(hdr); si = hdr.index("# Samples") - H; ii = hdr.index("Instructions Executed") - H
stall_cols = [k - H for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
lines = [(f, r) for f, r in lines if len(r) >= H and (r[si] or "0").isdigit()]
tot = sum(int(r[si] or 0) for _, r in lines)
toti = sum(int(r[ii] or 0) for _, r in lines)
print(f"== {kern}: {seen_fn} function table(s), total samples {tot}, warp instructions {toti}")
lines.sort(key=lambda fr: -int(fr[1][si] or 0))
for f, r in lines[:N]:
    st = sorted(((int(r[k] or 0), hdr[k + H][6:]) for k in stall_cols), reverse=True)[:3]
    st = " ".join(f"{n}:{c}" for c, n in st if c)
    print(f"{100*int(r[si] or 0)/max(tot,1):5.1f}% inst={100*int(r[ii] or 0)/max(toti,1):4.1f}% {f}:{r[0]:>4s} | {r[1].strip()[:100]} | {st}")
