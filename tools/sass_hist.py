"""Static SASS opcode histogram of the hot kernels (cuobjdump on the built objects; no GPU needed).
usage: python tools/sass_hist.py > profiles/r2_sass_hist.txt"""
import collections, glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOT = [  # (object glob, mangled-name regex, label)
    ("drice_encode_p0.o", r"encode_tile_kernelILi2ELi1ELb1ELi24ELi1E", "encode_tile_kernel<K=2, 1 CTA/SM, delta, 24 workers, table front-end>  (C2: M = 4)"),
    ("drice_encode_p0.o", r"encode_tile_kernelILi3ELi1ELb1ELi24ELi1E", "encode_tile_kernel<K=3, 1 CTA/SM, delta, 24 workers, table front-end>  (C3/C5: M = 8, L = 7000)"),
    ("drice_encode_p1.o", r"encode_tile_kernelILi4ELi2ELb1ELi12ELi0E", "encode_tile_kernel<K=4, arithmetic front-end>  (M = 16)"),
    ("drice_encode_p0.o", r"encode_long_pack_kernelILi3E", "encode_long_pack_kernel<K=3>  (few long waves)"),
    ("drice_decode.o", r"[0-9]parse_kernelILb1ELb0ELb0E", "parse_kernel<W<=10 bits, delta>  (C2)"),
    ("drice_decode.o", r"[0-9]parse_kernelILb0ELb0ELb0E", "parse_kernel<W>10 bits, delta>  (M = 8)"),
    ("drice_decode.o", r"scan_headers_kernel", "scan_headers_kernel"),
    ("drice_decode.o", r"rank_headers_kernel", "rank_headers_kernel"),
    ("drice_decode.o", r"[0-9]locate_kernelE", "locate_kernel (header chase: bulk async copies, SASS UBLKCP + SYNCS mbarrier)"),
    ("drice_decode.o", r"parse_wide_kernelILb0ELi512E", "parse_wide_kernel<delta>"),
    ("drice_decode.o", r"parse_long_kernelILb0E", "parse_long_kernel<delta>"),
]
KEY = ["UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "SHFL", "VOTE", "REDUX", "IMAD.WIDE", "VIADD.16x2", "PRMT", "LOP3", "SHF", "ATOMS", "ATOMG", "RED", "BAR", "NANOSLEEP", "LDL", "STL"]
for obj, rx, label in HOT:
    path = os.path.join(ROOT, "deltarice_b200", "csrc", "build", obj)
    names = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn = [m for m in re.findall(r"Function : (\S+)", names) if re.search(rx, m)]
    if not fn:
        print(f"== {label}: not found\n"); continue
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fn[0], path], capture_output=True, text=True).stdout
    ops = collections.Counter(); full = collections.Counter()
    for line in txt.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.x]+)", line)
        if m:
            full[m.group(1)] += 1
            ops[m.group(1).split(".")[0]] += 1
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", "-fun", fn[0], path], capture_output=True, text=True).stdout
    reg = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", res)
    n = sum(ops.values())
    print(f"== {label}")
    print(f"   {fn[0]}")
    print(f"   {n} instructions ({n * 16 / 1024:.1f} KB)" + (f", {reg.group(1)} registers, stack {reg.group(2)} B, static shared {reg.group(3)} B" if reg else ""))
    print("   " + "  ".join(f"{k} {v}" for k, v in ops.most_common(24)))
    keys = []
    for k in KEY:
        c = sum(v for kk, v in full.items() if kk == k or kk.startswith(k + ".") or (k in ("IMAD.WIDE", "VIADD.16x2") and kk.startswith(k)))
        if c: keys.append(f"{k} {c}")
    print("   key: " + "  ".join(keys) + "\n")
