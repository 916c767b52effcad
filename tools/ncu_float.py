#!/usr/bin/env python
"""Small fixed workload for an ncu capture of pdq_float_kernel: 1184 portrait 1024x768 RGB8 images (4 per CTA)."""
import sys
import torch
sys.path.insert(0, '/root/repo')
from rupphash_b200 import _lib, pdqhash
ctx = _lib.Context(0)
n = 1184
g = torch.Generator(device="cuda").manual_seed(3)
imgs = torch.empty((n, 1024, 768, 3), dtype=torch.uint8, device="cuda")
for s in range(0, n, 148):
    imgs[s:s + 148] = (torch.randn((148, 1024, 768, 3), generator=g, device="cuda") * 40 + 128).clamp_(0, 255).to(torch.uint8)
for _ in range(3):
    pdqhash.hash_batch(imgs, ctx=ctx)
print("ok", ctx.last_kernel_time())
