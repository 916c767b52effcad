#!/usr/bin/env python
"""Where the time of rh_hamming_group_multi goes (500k hashes, similarity 31) on all GPUs of the box:
wall time and the first GPU's event timeline, for the default / static-tiles / peer-copy modes."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from rupphash_b200 import _lib, scanner
from rupphash_b200.synth import planted_hashes
n = 500_000
hashes, low_conf = planted_hashes(n, seed=0xB200, n_clusters=5000, identical_block=1000, threshold=31)
_, h = bench.pinned(torch, hashes)
_, lc = bench.pinned(torch, low_conf)
out = {}
for name, nd, flags in (("1gpu", 1, 0), ("all_default_static_tiles", 0, 0), ("all_work_stealing", 0, 4), ("all_peercopy", 0, 1)):
    g = _lib.Group(n_dev=nd, flags=flags) if nd else _lib.Group(flags=flags)
    lab = torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    rows = []
    for rep in range(6):
        t0 = time.perf_counter()
        scanner.group_labels_multi(g, h, 31, low_conf=lc, out=lab)
        w = (time.perf_counter() - t0) * 1e3
        t = g.last_times(); t["python_wall_ms"] = w
        rows.append(t)
    out[name] = {k: float(np.median([r[k] for r in rows[2:]])) for k in rows[0]}
    out[name]["n_gpus"] = g.size
    print(name, json.dumps(out[name]), flush=True)
    g.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "group_timeline.json"), "w"), indent=1)
