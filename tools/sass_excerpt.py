#!/usr/bin/env python3
"""Hot part of an `ncu --page source --csv --print-source sass` export as a small text file: every
SASS instruction executed at least `floor` (default 2 %) as often as the most executed one, with its
execution count and lanes per execution -- the loops the DESIGN.md instruction counts refer to.
usage: python tools/sass_excerpt.py <export.csv> <out.txt> [floor]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
floor = float(sys.argv[3]) if len(sys.argv) > 3 else 0.02
hdr = rows[1]
ia, isrc, ie, it = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
base = int(rows[2][ia], 16)
R = [(int(r[ia], 16) - base, r[isrc].strip(), int(r[ie]), float(r[it])) for r in rows[2:]]
top = max(e for _, _, e, _ in R)
tot = sum(e for _, _, e, _ in R)
with open(sys.argv[2], "w") as out:
    out.write(f"# {rows[0][1] if len(rows[0]) > 1 else ''}\n# warp instructions {tot}, SASS lines {len(R)}; "
              f"shown: executed >= {floor:.0%} of the hottest line\n# offset  executed  lanes  instruction\n")
    gap = False
    for a, s, e, t in R:
        if e >= top * floor:
            if gap:
                out.write("  ...\n")
            out.write(f"{a:05x} {e:>11} {t:5.1f}  {s}\n")
            gap = False
        else:
            gap = True
