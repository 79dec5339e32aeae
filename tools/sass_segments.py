#!/usr/bin/env python3
"""Per-segment view of an `ncu --page source --csv --print-source sass` export: consecutive SASS
instructions with the same execution count and lane count are one segment (a basic block, in
practice); prints each segment's share of the kernel's warp instructions and its lanes per
instruction.   usage: python tools/sass_segments.py <export.csv> [min share, default 0.004]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
floor = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
hdr = rows[1]
ia, isrc, ie, it = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
base = int(rows[2][ia], 16)
R = [(int(r[ia], 16) - base, r[isrc].strip(), int(r[ie]), float(r[it])) for r in rows[2:]]
tot = sum(e for _, _, e, _ in R)
print(rows[0][1] if len(rows[0]) > 1 else "", "warp instructions", tot, "sass lines", len(R))


def op(s):
    f = s.split()
    return f[1] if f[0].startswith("@") else f[0]


seg, cur = [], None
for a, s, e, t in R:
    if cur and abs(e - cur["e"]) <= 0.02 * max(e, cur["e"], 1) and abs(t - cur["t"]) < 0.6:
        cur["n"] += 1
        cur["sum"] += e
        cur["ops"].append(op(s))
    else:
        if cur:
            seg.append(cur)
        cur = {"a": a, "e": e, "t": t, "n": 1, "sum": e, "ops": [op(s)]}
seg.append(cur)
MARK = ("TEX", "LDG", "LDS", "STS", "MUFU", "REDUX", "VOTE", "FRND", "F2I", "I2F", "WARPSYNC", "SHFL", "ATOM", "FLO", "STG")
for c in seg:
    if c["sum"] > tot * floor:
        key = [o for o in c["ops"] if o.startswith(MARK)]
        print(f"{c['a']:05x} n={c['n']:3d} exec={c['e']:>11} lanes={c['t']:4.1f} share={c['sum'] / tot:6.2%} {' '.join(key[:12])}")
