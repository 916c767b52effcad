import sys, json, numpy as np, torch
sys.path.insert(0, '/root/repo')
from rupphash_b200 import _lib, pdqhash
ctx = _lib.Context(0)
out = {}
for name, (h, w) in {"portrait_1024x768": (1024, 768), "portrait_512x384": (512, 384), "small_256x256": (256, 256), "landscape_768x1024": (768, 1024)}.items():
    n = 1024
    g = torch.Generator(device="cuda").manual_seed(3)
    imgs = (torch.randn((n, h, w, 3), generator=g, device="cuda") * 40 + 128).clamp_(0, 255).to(torch.uint8)
    pdqhash.hash_batch(imgs, ctx=ctx)
    ms = []
    for _ in range(5):
        pdqhash.hash_batch(imgs, ctx=ctx)
        ms.append(ctx.last_kernel_time()[0])
    t = float(np.median(ms)) * 1e-3
    out[name] = {"images_per_s": n / t, "ms_per_batch": t * 1e3}
    del imgs
print(json.dumps(out))
