#!/usr/bin/env python
"""Where the executed warp instructions and the stall samples of one kernel fall in the CUDA source: per 20-line region
and per line, launches of the report added up (needs -lineinfo + --import-source on).
usage: ncu_src_regions.py report.ncu-rep kernel_regex [N lines]"""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur=""; hdr=None; agg=collections.defaultdict(lambda:[0,0,""])
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if hdr and r[0]!="" and len(r)>=len(hdr):
        H=len(hdr); si=hdr.index("# Samples")-H; ii=hdr.index("Instructions Executed")-H
        try:
            a=agg[(cur,int(r[0]))]; a[0]+=int(r[si] or 0); a[1]+=int(r[ii] or 0); a[2]=r[1].strip()[:90]
        except ValueError: pass
ts=sum(a[0] for a in agg.values()); ti=sum(a[1] for a in agg.values())
print("total samples",ts,"instr",ti)
# by region of 25 lines
reg=collections.defaultdict(lambda:[0,0])
for (f,l),a in agg.items():
    k=(f,l//20*20); reg[k][0]+=a[0]; reg[k][1]+=a[1]
print("--- regions (file, line/20) by instructions")
for k,v in sorted(reg.items(), key=lambda kv:-kv[1][1])[:30]:
    print(f"{k[0]}:{k[1]:4d}-{k[1]+19:4d} inst {100*v[1]/ti:5.1f}% samples {100*v[0]/ts:5.1f}%")
print("--- lines by instructions")
for (f,l),a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:N]:
    print(f"inst {100*a[1]/ti:4.1f}% smp {100*a[0]/ts:4.1f}% {f}:{l:4d} | {a[2]}")
