"""Single very large / very elongated images (up to 108 MP) through rh_pdq_hash_batch and rh_phash_batch against the oracle."""
import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as orc
from rupphash_b200 import _lib, pdqhash, phash
ctx = _lib.Context(0)
rng = np.random.default_rng(5)
ok_all = True
for (h, w, ch) in [(9000, 12000, 3), (6000, 4000, 3), (4000, 6000, 4), (16384, 100, 3), (100, 16384, 1), (40000, 8, 1), (7, 30000, 3), (3333, 5000, 1)]:
    base = rng.integers(0, 256, size=(h // 16 + 1, w // 16 + 1, ch), dtype=np.uint8)
    img = np.repeat(np.repeat(base, 16, axis=0), 16, axis=1)[:h, :w].copy()
    img ^= rng.integers(0, 8, size=img.shape, dtype=np.uint8)
    arr = img[None] if ch > 1 else img[None, ..., 0]
    t0 = time.time()
    got = pdqhash.hash_batch(np.ascontiguousarray(arr), want_coeffs=True, want_dihedral=True, ctx=ctx)
    t1 = time.time()
    want = orc.pdq_batch(img[None], layout={3: 0, 4: 1, 1: 2}[ch], threads=1, want_coeffs=True, want_dihedral=True)
    ok = (np.array_equal(got["hash"], want["hash"]) and np.array_equal(got["coeffs"].view(np.uint32), want["coeffs"].view(np.uint32))
          and np.array_equal(got["quality"], want["quality"]) and np.array_equal(got["dihedral"], want["dihedral"])
          and np.array_equal(got["valid"], want["valid"]))
    pok = None
    if w * ch <= 16384:
        gp = phash.DctPhash(ctx).hash_batch(np.ascontiguousarray(arr))
        pok = int(gp[0]) == orc.phash_image(np.ascontiguousarray(arr[0]), layout={3: 0, 4: 1, 1: 2}[ch])[0]
    print((h, w, ch), "pdq", ok, "phash", pok, "gpu %.3fs" % (t1 - t0), flush=True)
    ok_all = ok_all and ok and (pok is not False)
print("ALL OK" if ok_all else "FAILED")
