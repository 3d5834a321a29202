"""Per-source-line instruction counts from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep kernel-regex [min_pct]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = None
items = []
for r in rows:
    if r and r[0] == "Line No":
        h = r
        iI, iS = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
        continue
    if h is None or len(r) <= iI or not r[0].strip():
        continue
    try:
        items.append((int(r[0]), r[1], int(r[iI]), int(r[iS] or 0)))
    except ValueError:
        pass
tot = sum(i[2] for i in items)
tst = sum(i[3] for i in items) or 1
print("total warp-instructions", tot)
for l, s, n, st in items:
    if n >= tot * minp / 100 or st >= tst * minp / 100:
        print(f"{l:4d} {n / tot * 100:6.2f}%  stall {st / tst * 100:5.1f}%  {s.strip()[:110]}")
