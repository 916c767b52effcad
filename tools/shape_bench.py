#!/usr/bin/env python
"""PDQ throughput by image shape on one GPU (device-resident images, kernel time from CUDA events):
which kernel serves each shape and what fraction of the HBM roofline (3 W H bytes per image over the
measured copy bandwidth) it reaches.  Writes gpurun_out/shape_bench.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rupphash_b200 import _lib, pdqhash  # noqa: E402

SHAPES = {   # name: (h, w, images, kernel that serves it)
    "landscape_768x1024": (768, 1024, 4096, "pdq_fused_kernel (integer passes 1-2)"),
    "portrait_1024x768": (1024, 768, 4096, "pdq_float_kernel"),
    "square_512x512": (512, 512, 8192, "pdq_fused_kernel"),
    "portrait_512x384": (512, 384, 8192, "pdq_float_kernel"),
    "landscape_384x512": (384, 512, 8192, "pdq_fused_kernel"),
    "small_256x256": (256, 256, 16384, "pdq_float_kernel"),
    "wide_180x512": (180, 512, 8192, "pdq_float_kernel (512 wide, below the fused kernel's height range)"),
    "odd_500x375": (500, 375, 4096, "generic pipeline (width not a multiple of 8)"),
}


def main():
    ctx = _lib.Context(0)
    hbm = bench.peaks()[0]
    out = {"hbm_gbs": hbm}
    for name, (h, w, n, kernel) in SHAPES.items():
        g = torch.Generator(device="cuda").manual_seed(3)
        imgs = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
        for s in range(0, n, 256):
            m = min(256, n - s)
            imgs[s:s + m] = (torch.randn((m, h, w, 3), generator=g, device="cuda") * 40 + 128).clamp_(0, 255).to(torch.uint8)
        pdqhash.hash_batch(imgs, ctx=ctx)
        ms = []
        for _ in range(7):
            pdqhash.hash_batch(imgs, ctx=ctx)
            ms.append(ctx.last_kernel_time()[0])
        t = float(np.median(ms)) * 1e-3
        rate = n / t
        out[name] = {"images_per_s": rate, "ms_per_batch": t * 1e3, "images": n, "kernel": kernel,
                     "hbm_roofline_frac": rate * (3 * h * w + 36) / 1e9 / hbm}
        print(name, out[name], flush=True)
        del imgs
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "shape_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
