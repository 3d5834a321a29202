"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, average and
share per kernel (cold-cache, serialised per-launch times: compare shares, not absolutes).
usage: python tools/ncu_launches.py launches.csv [title]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
iK, iV, iU, iG = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Grid Size")
acc = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[iV].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iU], 1.0)
    k = r[iK]
    if any(t in k for t in ("encode_", "parse_kernel", "locate_kernel")):
        k = k + "  grid " + r[iG].replace(" ", "")
    a = acc.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in acc.values())
ours = {k: a for k, a in acc.items() if "drice" in k or "encode_" in k or "parse_kernel" in k or "locate_kernel" in k or "word_copy" in k}
tot_ours = sum(a[1] for a in ours.values())
if len(sys.argv) > 2: print(sys.argv[2])
print("(cold-cache, serialised per-launch times: compare shares, not absolutes)")
print(f"all kernels: {sum(a[0] for a in acc.values())} launches, {tot / 1e3:.2f} ms; codec kernels: {sum(a[0] for a in ours.values())} launches, {tot_ours / 1e3:.2f} ms")
for k, a in sorted(acc.items(), key=lambda kv: -kv[1][1])[:14]:
    tag = "*" if k in ours else " "
    share = f"{a[1] / tot_ours * 100:5.1f}% of codec" if k in ours else ""
    print(f"{tag} {k[:110]:110s} launches {a[0]:5d}  avg {a[1] / a[0]:9.1f} us  {a[1] / tot * 100:5.1f}% of all  {share}")
