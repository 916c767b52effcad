#!/usr/bin/env python
"""Region totals of the fused PDQ kernel from an ncu cuda,sass source dump.
usage: ncu_regions.py dump.csv n_images"""
import contextlib
import io
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ncu_lines  # noqa: E402

dump, nimg = sys.argv[1], int(sys.argv[2])
fused = os.path.join(ROOT, "rupphash_b200/csrc/pdq_fused.cu")
tail = os.path.join(ROOT, "rupphash_b200/csrc/pdq_tail.cuh")
sys.argv = ["x", dump, fused, tail, "--top", "0"]
with contextlib.redirect_stdout(io.StringIO()):
    agg = ncu_lines.main()
tot = sum(a["smp"] for a in agg.values())
toti = sum(a["inst"] for a in agg.values())
src = open(fused).read().split("\n")


def find(s):
    return next(i + 1 for i, l in enumerate(src) if s in l)


marks = [("front_end", find("template <int BYTES>"), find("- edge columns ----")),
         ("edge", find("- edge columns ----"), find("- chain phase ----")),
         ("chain", find("- chain phase ----"), find("- tail ----")),
         ("pass4", find("- tail ----"), find("pdq_fused_kernel(const FusedArgs a)")),
         ("kernel body", find("pdq_fused_kernel(const FusedArgs a)"), len(src))]
keys = ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_math", "stall_selected")
for name, a, b in marks + [("pdq_tail.cuh", 0, 0), ("other", 0, 0)]:
    if b:
        sel = [v for (o, l, t), v in agg.items() if o == "pdq_fused.cu" and a <= l < b]
    else:
        sel = [v for (o, l, t), v in agg.items() if o == name]
    s = sum(v["smp"] for v in sel)
    i = sum(v["inst"] for v in sel)
    st = " ".join(f"{k[6:]}={100 * sum(v[k] for v in sel) / tot:4.1f}" for k in keys)
    print(f"{name:12s} smp {100 * s / tot:5.1f}%  inst {100 * i / toti:5.1f}% = {i / nimg:8.0f} warp-inst/img | {st}")
print(f"total {toti / nimg:.0f} warp-inst/img, {tot:.0f} samples")
