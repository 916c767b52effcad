#!/usr/bin/env python
"""Times the Hamming tile kernel on ONE GPU: the whole 500k search, each rank's share of an 8-rank
static split (what one GPU of eight does), the kernel variants, and the reference's default setting
(similarity 40, 8 variants).  Writes gpurun_out/hamming_probe.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from rupphash_b200 import _lib, scanner  # noqa: E402
from rupphash_b200.synth import planted_hashes, random_variants  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
    ctx = _lib.Context(0)
    pk = ctx.measure_peaks()
    hashes, low_conf = planted_hashes(n, seed=0xB200, n_clusters=5000, identical_block=1000, threshold=31)
    d_h = torch.from_numpy(hashes).cuda()
    d_l = torch.from_numpy(low_conf).cuda()
    pairs = n * (n - 1) // 2
    out = {"n": n, "pairs": pairs, "peaks": pk}

    def timed(fn, reps=3):
        fn()
        ms, wall = [], []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            wall.append((time.perf_counter() - t0) * 1e3)
            ms.append(ctx.last_kernel_time()[0])
        return float(np.median(ms)), float(np.median(wall))

    for pf in (3, 0, 4, 5, 6, 7):
        ctx.set_option("hamming.prefilter", pf)
        k, w = timed(lambda: scanner.group_labels(d_h, 31, low_conf=d_l, ctx=ctx))
        popc = {3: 2, 4: 3, 0: 4, 5: 2, 6: 2, 7: 3}[pf]
        out[f"full_pf{pf}"] = {"tile_ms": k, "wall_ms": w, "pairs_per_s": pairs / (k * 1e-3),
                               "executed_popc_frac": pairs / (k * 1e-3) * popc / pk["popc_per_s"]}
    for pf in (4, 5, 6):     # the reference's default threshold: exact 128-bit prefix (3 POPC) vs OR bound (2 POPC)
        ctx.set_option("hamming.prefilter", pf)
        k, w = timed(lambda: scanner.group_labels(d_h, 40, low_conf=d_l, ctx=ctx))
        popc = {4: 3, 5: 2, 6: 2}[pf]
        out[f"sim40_pf{pf}"] = {"tile_ms": k, "wall_ms": w, "pairs_per_s": pairs / (k * 1e-3),
                                "executed_popc_frac": pairs / (k * 1e-3) * popc / pk["popc_per_s"]}
    for pf in (0, 7):        # the largest threshold the reference accepts
        ctx.set_option("hamming.prefilter", pf)
        k, w = timed(lambda: scanner.group_labels(d_h, 63, low_conf=d_l, ctx=ctx))
        out[f"sim63_pf{pf}"] = {"tile_ms": k, "wall_ms": w, "pairs_per_s": pairs / (k * 1e-3),
                                "executed_popc_frac": pairs / (k * 1e-3) * {0: 4, 7: 3}[pf] / pk["popc_per_s"]}
    ctx.set_option("hamming.prefilter", -1)
    for sim in (0, 8, 12, 16, 20, 24, 28, 31, 40, 48, 56, 63):
        k, w = timed(lambda: scanner.group_labels(d_h, sim, low_conf=d_l, ctx=ctx), reps=2)
        out[f"auto_sim{sim}"] = {"tile_ms": k, "wall_ms": w, "pairs_per_s": pairs / (k * 1e-3), "variant": ctx.hamming_last_variant()}
    shards = []
    for r in range(8 if len(sys.argv) <= 2 else 0):
        k, w = timed(lambda: scanner.group_shard(d_h, 31, r, 8, low_conf=d_l, ctx=ctx), reps=2)
        shards.append({"rank": r, "tile_ms": k, "wall_ms": w})
    out["shards_of_8"] = shards
    out["shard_sum_ms"] = sum(s["tile_ms"] for s in shards)
    out["shard_max_ms"] = max([s["tile_ms"] for s in shards] or [0.0])
    # the reference's default: similarity 40 with 8 variants per file
    m = min(n, 200_000)
    var = random_variants(hashes[:m], seed=5)
    d_v = torch.from_numpy(var).cuda()
    for sim in (31, 40, 63):
        k, w = timed(lambda: scanner.group_labels(d_h[:m], sim, variants=d_v, low_conf=d_l[:m], ctx=ctx), reps=2)
        pr = 8 * m * (m - 1) // 2
        out[f"variants8_sim{sim}"] = {"n": m, "tile_ms": k, "wall_ms": w, "pairs_per_s": pr / (k * 1e-3),
                                      "variant": ctx.hamming_last_variant()}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "hamming_probe.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
