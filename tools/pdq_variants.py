#!/usr/bin/env python
"""A-B timing of the fused PDQ kernel's switches on one GPU (device-resident 1024x768 RGB8 pool, the
bench shape, plus 512x512): pdq.variant bits, prefetch distance.  Checks every variant's hashes against
the default's.  Writes gpurun_out/pdq_variants.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from rupphash_b200 import _lib  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4, 7]
    ctx = _lib.Context(0)
    L = _lib.lib()
    hbm = bench.peaks()[0]
    pool = bench.synth_pool_device(torch, n, seed=0xB200)
    out_hash = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
    out_q = torch.empty((n,), dtype=torch.float32, device="cuda")
    res = {"n": n, "hbm_gbs": hbm}

    def run(px, w, h, reps=6):
        ms = []
        for _ in range(reps):
            ctx.check(L.rh_pdq_hash_batch(ctx.handle, px.data_ptr(), _lib.LAYOUT_RGB8, px.shape[0], w, h, 0, 0,
                                          out_hash.data_ptr(), out_q.data_ptr(), None, None, None))
            ms.append(ctx.last_kernel_time()[0])
        return float(np.median(ms[2:]))

    # interleaved A-B: every round runs each variant once, so that clock / thermal drift hits all alike
    ref = None
    samples = {v: [] for v in variants}
    same = {}
    for rnd in range(24):
        for v in variants:
            ctx.set_option("pdq.variant", v)
            ctx.check(L.rh_pdq_hash_batch(ctx.handle, pool.data_ptr(), _lib.LAYOUT_RGB8, n, 1024, 768, 0, 0,
                                          out_hash.data_ptr(), out_q.data_ptr(), None, None, None))
            if rnd >= 4:
                samples[v].append(ctx.last_kernel_time()[0])
            if rnd == 0:
                h = out_hash.cpu().numpy().copy()
                if ref is None:
                    ref = h
                same[v] = bool(np.array_equal(h, ref))
    for v in variants:
        ms, best = float(np.median(samples[v])), float(np.min(samples[v]))
        rate = n / (ms * 1e-3)
        res[f"variant{v}"] = {"ms_median": ms, "ms_min": best, "img_per_s": rate,
                              "frac": rate * bench.ALGO_BYTES_PER_IMAGE / 1e9 / hbm, "same_hashes": same[v]}
        print(v, res[f"variant{v}"], flush=True)
    ctx.set_option("pdq.variant", 0)
    for rows in (16, 32, 48, 64, 96, 16):
        ctx.set_option("pdq.prefetch_rows", rows)
        ms = run(pool, 1024, 768)
        res[f"prefetch_rows{rows}"] = {"ms": ms, "img_per_s": n / (ms * 1e-3)}
        print("pf rows", rows, ms, flush=True)
    ctx.set_option("pdq.prefetch_rows", 32)
    # 512 x 512 (configs[3] shape): 3 x as many images in the same bytes
    sq = pool.reshape(-1)[: (n * 3) * 512 * 512 * 3].reshape(n * 3, 512, 512, 3)
    out_hash = torch.empty((n * 3, 32), dtype=torch.uint8, device="cuda")
    out_q = torch.empty((n * 3,), dtype=torch.float32, device="cuda")
    ms = run(sq, 512, 512)
    rate = 3 * n / (ms * 1e-3)
    res["square512"] = {"ms": ms, "img_per_s": rate, "frac": rate * (512 * 512 * 3 + 36) / 1e9 / hbm}
    print("512x512", res["square512"], flush=True)
    for v in (8, 0, 8, 0):     # the per-SM turn lock on a shape whose front end is a small share
        ctx.set_option("pdq.variant", v)
        ms = run(sq, 512, 512)
        res.setdefault(f"square512_variant{v}", []).append({"ms": ms, "img_per_s": 3 * n / (ms * 1e-3)})
        print("512x512 variant", v, ms, flush=True)
    ctx.set_option("pdq.variant", 0)
    ctx.set_option("pdq.phase_clocks", 1)
    run(pool, 1024, 768, reps=3)
    run(sq, 512, 512, reps=3)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "pdq_variants.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
