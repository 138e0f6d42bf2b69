#!/usr/bin/env python
"""Top stall sites of one kernel from an ncu report: python profiles/stalls.py <rep> <kernel regex> [launch#]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
nth = sys.argv[3] if len(sys.argv) > 3 else "1"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{pat}:{nth}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iS]) for r in data)
print(rows[0][1][:100]); print("total samples", tot, "instructions", sum(int(r[iEx]) for r in data))
agg = {s: sum(int(r[hdr.index(s)]) for r in data) for s in stalls}
print([(k, round(v / tot, 3)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]])
for r in sorted(data, key=lambda r: -int(r[iS]))[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]:
    st = {s: int(r[hdr.index(s)]) for s in stalls if int(r[hdr.index(s)]) > 0}
    print(f"{int(r[iS]):6d} {int(r[iS])/tot:6.1%} ex={r[iEx]:>8} {r[iSrc].strip()[:64]:64s}", sorted(st.items(), key=lambda kv: -kv[1])[:2])
