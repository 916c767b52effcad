import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from rupphash_b200 import _lib, pdqhash
ctx = _lib.Context(0)
for (h, w, n) in ((1024, 768, 4096), (512, 384, 4096), (256, 256, 4096)):
    g = torch.Generator(device="cuda").manual_seed(3)
    imgs = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
    for s in range(0, n, 256):
        imgs[s:s+256] = (torch.randn((256, h, w, 3), generator=g, device="cuda") * 40 + 128).clamp_(0, 255).to(torch.uint8)
    pdqhash.hash_batch(imgs, ctx=ctx)
    ctx.set_option("pdq.phase_clocks", 1)
    pdqhash.hash_batch(imgs, ctx=ctx)
    ctx.set_option("pdq.phase_clocks", 0)
    ms = []
    for _ in range(5):
        pdqhash.hash_batch(imgs, ctx=ctx); ms.append(ctx.last_kernel_time()[0])
    print(h, w, n / (np.median(ms) * 1e-3), flush=True)
    del imgs
