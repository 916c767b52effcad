#!/usr/bin/env python
"""Executed-opcode mix per source region from an `ncu --page source --csv --print-source=cuda,sass` dump.
usage: ncu_opmix.py dump.csv n_images file.cu first_line last_line [first last ...]"""
import collections
import csv
import re
import sys

dump, nimg, src = sys.argv[1], int(sys.argv[2]), sys.argv[3]
ranges = [(int(a), int(b)) for a, b in zip(sys.argv[4::2], sys.argv[5::2])]
text = [l.rstrip("\n").strip() for l in open(src)]
rows = list(csv.reader(open(dump)))
hdr = next(r for r in rows if r and r[0] == "Line No")
iE = hdr.index("Instructions Executed")
cur = None
mix = {r: collections.Counter() for r in ranges}
for r in rows:
    if len(r) < len(hdr):
        continue
    if r[0].isdigit():       # a CUDA source line: following SASS rows belong to it
        ln = int(r[0])
        cur = ln if 0 < ln <= len(text) and text[ln - 1] == r[1].strip() else None
        continue
    if r[2] in ("-", "") or cur is None:
        continue
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[3])
    if not m:
        continue
    try:
        n = int(r[iE])
    except ValueError:
        continue
    for rg in ranges:
        if rg[0] <= cur < rg[1]:
            mix[rg][m.group(2)] += n
for rg, c in mix.items():
    tot = sum(c.values())
    print(f"lines {rg[0]}-{rg[1]}: {tot / nimg:.0f} warp-inst/img")
    for op, n in c.most_common(28):
        print(f"   {op:26s} {n / nimg:9.0f} {100 * n / tot:5.1f}%")
