import csv, subprocess, sys
rep, pat, skip = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{pat}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
out, cur, hdr = [], None, None
for r in rows:
    if not r: continue
    if r[0] in ("File Path","File Name"):
        cur, hdr = r[1].split("/")[-1], None; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and "Instructions Executed" in hdr:
        iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try: out.append((cur, int(r[0]), r[1].strip()[:110], int(r[iE] or 0), int(r[iS] or 0)))
        except ValueError: pass
tot = sum(o[3] for o in out); smp = sum(o[4] for o in out)
print("total warp-instructions", tot, "samples", smp)
for o in sorted(out, key=lambda o: -o[int(__import__("os").environ.get("SORTCOL","4"))])[:top]:
    print(f"{o[3]:9d} {o[3]/max(tot,1):6.1%} smp={o[4]/max(smp,1):6.1%} {o[0]}:{o[1]:<4d} {o[2]}")
