#!/usr/bin/env python
"""Executed warp-instructions and stall samples per CUDA source line:
   python profiles/srclines.py <rep> <kernel regex> [launch#] [top]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
nth = sys.argv[3] if len(sys.argv) > 3 else "1"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id",
                      f"::regex:{pat}:{nth}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
out, cur, hdr = [], None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        out.append((cur, int(r[0]), r[1].strip()[:100], int(r[iE]), int(r[iS])))
tot = sum(o[3] for o in out)
smp = sum(o[4] for o in out)
print("total warp-instructions", tot, "samples", smp)
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{o[3]:9d} {o[3]/tot:6.1%} smp={o[4]/max(smp,1):6.1%} {o[0]}:{o[1]:<4d} {o[2]}")
