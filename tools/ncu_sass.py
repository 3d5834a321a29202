"""Hot SASS instructions of one kernel launch from an .ncu-rep (--page source).
usage: python tools/ncu_sass.py report.ncu-rep kernel-regex [min_pct] [launch-index]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
blocks, cur, h = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append((r[1], cur))
    elif r and r[0] == "Address":
        h = r
    elif cur is not None and h and len(r) == len(h):
        cur.append(r)
name, rs = blocks[which]
iI, iS, iT = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)"), h.index("Avg. Threads Executed")
iW, iWi = h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Ideal")
tot = sum(int(r[iI]) for r in rs)
tst = sum(int(r[iS]) for r in rs) or 1
print(name, "total warp-instructions", tot, "stall samples", tst)
for r in rs:
    n, st = int(r[iI]), int(r[iS])
    if n >= tot * minp / 100 or st >= tst * minp / 100:
        wf = f" wf {r[iW]}/{r[iWi]}" if r[iW] not in ("0", "") else ""
        print(f"{r[0][-5:]} {n / tot * 100:5.2f}% st {st / tst * 100:5.2f}% thr {r[iT]:>3s}  {r[1].strip()[:100]}{wf}")
