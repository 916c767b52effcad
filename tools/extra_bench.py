#!/usr/bin/env python
"""Secondary measurements (not the headline bench line): the other BASELINE shapes.
  * config 4 hashing shape: 512x512 RGB8, hash + quality + 256 coefficients + 8 dihedral hashes
  * config 4 grouping shape: 8 dihedral query variants per file at threshold 31
  * config 5 grouping scale on one GPU: 1M hashes, without and with 8 variants per file
Prints one JSON object; run on a GPU box."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rupphash_b200 import _lib, pdqhash, scanner  # noqa: E402
from rupphash_b200.synth import planted_hashes  # noqa: E402


def main():
    ctx = _lib.Context(0)
    out = {}
    # ---- 512 x 512, all outputs
    n = 8192
    g = torch.Generator(device="cuda").manual_seed(3)
    imgs = (torch.randn((n, 512, 512, 3), generator=g, device="cuda") * 40 + 128).clamp_(0, 255).to(torch.uint8)
    for want_all in (False, True):
        pdqhash.hash_batch(imgs, want_coeffs=want_all, want_dihedral=want_all, ctx=ctx)
        ms = []
        for _ in range(5):
            pdqhash.hash_batch(imgs, want_coeffs=want_all, want_dihedral=want_all, ctx=ctx)
            ms.append(ctx.last_kernel_time()[0])
        t = float(np.median(ms)) * 1e-3
        bytes_per = 512 * 512 * 3 + 36 + (1024 + 224 if want_all else 0)
        out["pdq_512x512_all_outputs" if want_all else "pdq_512x512_hash_only"] = {
            "images_per_s": n / t, "ms_per_batch": t * 1e3, "batch": n,
            "hbm_frac_of_6557": n * bytes_per / t / 6557.4e9}
    del imgs
    # ---- grouping with 8 variants per file
    nh = 250_000
    hashes, low_conf = planted_hashes(nh, seed=5, n_clusters=2500, identical_block=500)
    rng = np.random.default_rng(1)
    variants = rng.integers(0, 256, size=(nh, 8, 32), dtype=np.uint8)
    variants[:, 0] = hashes
    dh, dv, dl = (torch.from_numpy(x).cuda() for x in (hashes, variants, low_conf))
    scanner.group_labels(dh, 31, variants=dv, low_conf=dl, ctx=ctx)
    ms = []
    for _ in range(3):
        t0 = time.perf_counter()
        labels, cnt = scanner.group_labels(dh, 31, variants=dv, low_conf=dl, ctx=ctx)
        torch.cuda.synchronize()
        ms.append((time.perf_counter() - t0, ctx.last_kernel_time()[0]))
    wall, kms = min(ms)
    pairs = 8 * nh * (nh - 1) // 2
    out["hamming_250k_x8_variants"] = {"pairs": pairs, "pairs_per_s": pairs / wall, "wall_ms": wall * 1e3,
                                       "tile_kernel_ms": kms, "edges": cnt}
    del dh, dv, dl
    # ---- config 5 grouping scale on ONE GPU: 1M hashes, own hash only and with 8 variants per file
    nh = 1_000_000
    hashes, low_conf = planted_hashes(nh, seed=9, n_clusters=10000, identical_block=500)
    dh, dl = torch.from_numpy(hashes).cuda(), torch.from_numpy(low_conf).cuda()
    for name, dv in (("hamming_1M", None), ("hamming_1M_x8_variants", "make")):
        if dv == "make":
            dv = torch.randint(0, 256, (nh, 8, 32), dtype=torch.uint8, device="cuda",
                               generator=torch.Generator(device="cuda").manual_seed(2))
            dv[:, 0] = dh
        wall = 1e9
        for _ in range(2):   # the first call also grows the library's scratch buffers
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            labels, cnt = scanner.group_labels(dh, 31, variants=dv, low_conf=dl, ctx=ctx)
            torch.cuda.synchronize()
            wall = min(wall, time.perf_counter() - t0)
        pairs = (8 if dv is not None else 1) * nh * (nh - 1) // 2
        out[name] = {"pairs": pairs, "pairs_per_s": pairs / wall, "wall_ms": wall * 1e3,
                     "tile_kernel_ms": ctx.last_kernel_time()[0], "edges": cnt,
                     "groups": int((torch.bincount(labels.long(), minlength=nh) > 1).sum())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
