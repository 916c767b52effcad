#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source=cuda,sass` dump per source line / region.
usage: ncu_lines.py dump.csv file1.cu [file2.cuh ...] [--top N]"""
import collections
import csv
import sys


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    top = 25
    if "--top" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1])
        args = [a for a in args if a != str(top)]
    dump, files = args[0], args[1:]
    texts = {f: [l.rstrip("\n") for l in open(f)] for f in files}
    rows = list(csv.reader(open(dump)))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
    cols = {k: hdr.index(k) for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math",
                                      "stall_mio", "stall_not_selected", "stall_selected")}

    def fl(x):
        try:
            return float(x)
        except ValueError:
            return 0.0

    agg = collections.defaultdict(lambda: collections.Counter())
    for r in rows:
        if len(r) < len(hdr):
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        owner = "other"
        for f, t in texts.items():
            if 0 < ln <= len(t) and t[ln - 1].strip() == r[1].strip() and r[1].strip():
                owner = f.split("/")[-1]
                break
        a = agg[(owner, ln, r[1].strip()[:80])]
        a["smp"] += fl(r[iS])
        a["inst"] += fl(r[iI])
        for k, i in cols.items():
            a[k] += fl(r[i])
    tot = sum(a["smp"] for a in agg.values()) or 1
    toti = sum(a["inst"] for a in agg.values()) or 1
    print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
    byfile = collections.defaultdict(lambda: collections.Counter())
    for (o, ln, t), a in agg.items():
        byfile[o].update(a)
    for o, a in byfile.items():
        print(f"  {o:22s} samples {100 * a['smp'] / tot:5.1f}%  inst {100 * a['inst'] / toti:5.1f}%")
    print("top lines by samples:")
    for (o, ln, t), a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:top]:
        st = " ".join(f"{k[6:]}={a[k]:.0f}" for k in cols if a[k] > 0.02 * a["smp"])
        print(f"  {o[:14]:14s}:{ln:4d} smp {100 * a['smp'] / tot:5.1f}% inst {100 * a['inst'] / toti:5.1f}% [{st}] {t}")
    print("top lines by instructions:")
    for (o, ln, t), a in sorted(agg.items(), key=lambda kv: -kv[1]["inst"])[:top]:
        print(f"  {o[:14]:14s}:{ln:4d} inst {100 * a['inst'] / toti:5.1f}% smp {100 * a['smp'] / tot:5.1f}% {t}")
    return agg


if __name__ == "__main__":
    main()
