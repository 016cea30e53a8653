#!/usr/bin/env python3
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv | ncu_lines.py [top]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
cur_file = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] in ("Function Name",): continue
    if len(r) < len(hdr) or r[2] != "-": continue      # only CUDA-line summary rows (Address '-')
    try:
        ie = int(r[hdr.index("Instructions Executed")]); ns = int(r[hdr.index("# Samples")])
    except ValueError: continue
    out.append((ie, ns, cur_file, r[0], r[1].strip()[:100]))
ti = sum(o[0] for o in out); ts = sum(o[1] for o in out)
print(f"total warp-instructions {ti}, stall samples {ts}")
key = (lambda x: -x[0]) if (len(sys.argv) > 2 and sys.argv[2] == "inst") else (lambda x: -x[1])
for o in sorted(out, key=key)[:top]:
    print(f"{100*o[1]/max(ts,1):5.1f}% smp {100*o[0]/max(ti,1):5.1f}% inst  {o[2]}:{o[3]:>4}  {o[4]}")
