"""Time-bounded randomised parity run of both hot paths through the C ABI against the CPU oracle.

    python tools/fuzz_parity.py [--seconds 240] [--seed 1] [--out gpurun_out/fuzz.json]

The fixed cases live in tests/; this tool walks the input space the tests sample: random image
sizes around every kernel's domain borders (landscape fused kernel, float-chain kernel, generic
pipeline, general Box pre-downsample), all three layouts, padded rows / images, image content
from flat to pure noise, and for the grouping path random n, thresholds 0..63, variant counts,
has_hash / low_conf masks and hash populations (uniform, planted clusters, shared prefixes, blocks
of identical hashes).  Every case is compared bit for bit (hash, quality, coefficient bit
patterns, dihedral hashes; labels, edge count, edge multiset).  A mismatch prints the case's
parameters (they are all derived from --seed and the case number) and the run exits 1.
RH_B200_LIB=rupphash_b200/librupphash_b200_dbg.so runs the same cases on the build with in-kernel
index assertions (make -C rupphash_b200/csrc debug).
The oracle is the checker here (test infrastructure); nothing in the product path uses it.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import oracle as orc                                    # noqa: E402  (checker)
from rupphash_b200 import _lib, scanner                 # noqa: E402
from rupphash_b200.synth import planted_hashes, random_variants, synth_images  # noqa: E402


def pick_shape(rng):
    """(h, w) biased towards the borders between the three PDQ kernels"""
    kind = rng.integers(0, 10)
    if kind == 0:      # landscape fused kernel, no pre-downsample: plane 512 wide
        return int(rng.integers(190, 515)), 512
    if kind == 1:      # landscape fused kernel behind the exact 2x pre-downsample
        return int(rng.integers(190, 515)) * 2, 1024
    if kind == 2:      # float-chain kernel: widths 72..512 multiple of 8, heights 65..512
        return int(rng.integers(60, 516)), int(rng.integers(8, 66)) * 8
    if kind == 3:      # the same behind the 2x pre-downsample
        return int(rng.integers(258, 516)) * 2, int(rng.integers(33, 66)) * 16
    if kind == 4:      # widths just off the multiples of 8 / 16
        return int(rng.integers(60, 1030)), int(rng.integers(60, 1030))
    if kind == 5:      # general Box pre-downsample (one side above 1024 or odd)
        return int(rng.integers(500, 1700)), int(rng.integers(500, 1700))
    if kind == 6:      # tiny and degenerate
        return int(rng.integers(1, 80)), int(rng.integers(1, 80))
    if kind == 7:      # extreme aspect ratios
        a, b = int(rng.integers(5, 64)), int(rng.integers(600, 2600))
        return (a, b) if rng.integers(0, 2) else (b, a)
    if kind == 8:      # the headline shapes
        return [(768, 1024), (1024, 768), (512, 512), (384, 512), (512, 384), (256, 256)][int(rng.integers(0, 6))]
    return int(rng.integers(64, 520)), int(rng.integers(64, 520))


def make_images(rng, n, h, w, ch):
    kind = int(rng.integers(0, 7))
    if kind == 0:
        imgs = synth_images(n, h, w, seed=int(rng.integers(1, 1 << 30)), channels=ch)
    elif kind == 1:
        imgs = rng.integers(0, 256, size=(n, h, w, ch), dtype=np.uint8)
    elif kind == 2:    # flat images, all values
        imgs = np.empty((n, h, w, ch), np.uint8)
        for k in range(n):
            imgs[k] = rng.integers(0, 256, size=(1, 1, ch), dtype=np.uint8)
    elif kind == 3:    # saturated blocks: 0 / 255 only
        blk = int(rng.integers(1, 40))
        yy = (np.arange(h) // blk)[:, None]
        xx = (np.arange(w) // blk)[None, :]
        imgs = np.empty((n, h, w, ch), np.uint8)
        for k in range(n):
            t = rng.integers(0, 2, size=(h // blk + 1, w // blk + 1), dtype=np.uint8) * 255
            imgs[k] = t[yy, xx][..., None]
    elif kind == 4:    # smooth gradients (tie-heavy medians, low quality)
        gy = np.linspace(0, float(rng.integers(1, 256)), h)[:, None]
        gx = np.linspace(0, float(rng.integers(1, 256)), w)[None, :]
        base = np.clip(gy + gx, 0, 255).astype(np.uint8)
        imgs = np.repeat(np.repeat(base[None, :, :, None], n, axis=0), ch, axis=3).copy()
        imgs[:, :: max(1, h // 7)] ^= 1
    elif kind == 5:    # sparse bright pixels on black
        imgs = np.zeros((n, h, w, ch), np.uint8)
        m = rng.random((n, h, w)) < 0.01
        imgs[m] = 255
    else:              # noise with a strong channel imbalance
        imgs = rng.integers(0, 256, size=(n, h, w, ch), dtype=np.uint8)
        if ch >= 3:
            imgs[..., int(rng.integers(0, 3))] = int(rng.integers(0, 256))
    return imgs


def pdq_case(ctx, rng, case):
    import torch
    h, w = pick_shape(rng)
    ch = [3, 3, 4, 1][int(rng.integers(0, 4))]
    layout = {3: 0, 4: 1, 1: 2}[ch]
    n = int(rng.integers(1, 6)) if h * w > 300_000 else int(rng.integers(1, 14))
    imgs = make_images(rng, n, h, w, ch)
    pad_kind = int(rng.integers(0, 5))      # packed, packed, +16k, odd, rows rounded up to 16 bytes
    pad_row = [0, 0, 16 * int(rng.integers(1, 9)), int(rng.integers(1, 40)),
               (-w * ch) % 16 + 16 * int(rng.integers(0, 3))][pad_kind]
    pad_img = [0, 0, 16 * int(rng.integers(0, 300)), int(rng.integers(0, 500)), 16 * int(rng.integers(0, 9))][pad_kind]
    device_resident = bool(rng.integers(0, 2))
    desc = dict(case=case, path="pdq", h=h, w=w, ch=ch, n=n, pad_row=pad_row, pad_img=pad_img,
                device_resident=device_resident)
    want = orc.pdq_batch(imgs, layout=layout, threads=16, want_coeffs=True, want_dihedral=True)
    row_pitch = w * ch + pad_row
    img_pitch = h * row_pitch + pad_img
    buf = np.full((n * img_pitch,), 0x5A, np.uint8)
    for k in range(n):
        rows = buf[k * img_pitch: k * img_pitch + h * row_pitch].reshape(h, row_pitch)
        rows[:, : w * ch] = imgs[k].reshape(h, w * ch)
    got = {"hash": np.zeros((n, 32), np.uint8), "quality": np.zeros(n, np.float32),
           "coeffs": np.zeros((n, 256), np.float32), "dihedral": np.zeros((n, 8, 32), np.uint8),
           "valid": np.zeros(n, np.uint8)}
    src = torch.from_numpy(buf).cuda() if device_resident else buf
    ctx.check(_lib.lib().rh_pdq_hash_batch(ctx.handle, _lib.ptr(src), layout, n, w, h, row_pitch, img_pitch,
                                           _lib.ptr(got["hash"]), _lib.ptr(got["quality"]), _lib.ptr(got["coeffs"]),
                                           _lib.ptr(got["dihedral"]), _lib.ptr(got["valid"])))
    ok = (np.array_equal(got["valid"], want["valid"]) and np.array_equal(got["hash"], want["hash"])
          and np.array_equal(got["quality"].view(np.uint32), want["quality"].view(np.uint32))
          and np.array_equal(got["coeffs"].view(np.uint32), want["coeffs"].view(np.uint32))
          and np.array_equal(got["dihedral"], want["dihedral"]))
    return ok, desc


def make_hashes(rng, n, similarity):
    kind = int(rng.integers(0, 5))
    seed = int(rng.integers(1, 1 << 30))
    if kind == 0 or n < 64:
        return rng.integers(0, 256, size=(n, 32), dtype=np.uint8), np.zeros(n, np.uint8), "uniform"
    if kind == 1:
        h, lc = planted_hashes(n, seed=seed, threshold=max(1, min(similarity, 100)))
        return h, lc, "planted"
    if kind == 2:      # a shared prefix of 96..160 bits: the prefilter passes everything
        h, lc = planted_hashes(n, seed=seed, threshold=max(1, min(similarity, 100)))
        k = int(rng.integers(12, 21))
        h[:, :k] = rng.integers(0, 256, size=k, dtype=np.uint8)
        return h, lc, "shared_prefix_%d" % (8 * k)
    if kind == 3:      # few distinct values: big identical blocks
        base = rng.integers(0, 256, size=(int(rng.integers(1, 6)), 32), dtype=np.uint8)
        h = base[rng.integers(0, len(base), size=n)]
        return np.ascontiguousarray(h), (rng.random(n) < 0.3).astype(np.uint8), "few_values"
    # a dense ball: everything within a few bits of one centre
    centre = rng.integers(0, 256, size=32, dtype=np.uint8)
    h = np.repeat(centre[None], n, axis=0)
    flips = rng.integers(0, 256, size=(n, int(rng.integers(1, 30))))
    for j in range(flips.shape[1]):
        h[np.arange(n), flips[:, j] >> 3] ^= (1 << (flips[:, j] & 7)).astype(np.uint8)
    return h, np.zeros(n, np.uint8), "dense_ball"


def hamming_case(ctx, rng, case):
    n = int([rng.integers(1, 70), rng.integers(70, 3000), rng.integers(3000, 12000)][int(rng.integers(0, 3))])
    similarity = int([rng.integers(0, 32), rng.integers(32, 64), 63, 31, 40][int(rng.integers(0, 5))])
    hashes, low_conf, pop = make_hashes(rng, n, similarity)
    if pop in ("few_values", "dense_ball") and n > 4000:
        n = 4000                       # O(n^2) edges: keep the brute-force oracle in seconds
        hashes, low_conf = hashes[:n], low_conf[:n]
    use_variants = bool(rng.integers(0, 2))
    variants = n_variants = None
    if use_variants:
        variants = random_variants(hashes, seed=int(rng.integers(1, 1 << 30)))
        if rng.integers(0, 2):
            n_variants = rng.integers(0, 9, size=n).astype(np.uint8)
    has_hash = (rng.random(n) < 0.9).astype(np.uint8) if rng.integers(0, 3) == 0 else None
    lc = low_conf if rng.integers(0, 2) else None
    desc = dict(case=case, path="hamming", n=n, similarity=similarity, population=pop, variants=use_variants,
                n_variants=n_variants is not None, has_hash=has_hash is not None, low_conf=lc is not None)
    cap = 1 << 22
    ref_labels, ref_cnt, ref_edges = orc.group_generic(hashes, similarity, has_hash=has_hash, variants=variants,
                                                       n_variants=n_variants, low_conf=lc, use_mih=False, edges_cap=cap)
    labels, cnt = scanner.group_labels(hashes, similarity, has_hash=has_hash, variants=variants, n_variants=n_variants,
                                       low_conf=lc, ctx=ctx)
    ok = cnt == ref_cnt and np.array_equal(labels, ref_labels)
    if ok and ref_cnt <= cap:
        got_edges, cnt2 = scanner.edges(hashes, similarity, has_hash=has_hash, variants=variants,
                                        n_variants=n_variants, low_conf=lc, cap=cap, ctx=ctx)
        key = lambda e: np.sort(e[:, 0].astype(np.uint64) << np.uint64(32) | e[:, 1].astype(np.uint64))
        ok = cnt2 == ref_cnt and np.array_equal(key(got_edges), key(ref_edges))
    desc["edges"] = int(ref_cnt)
    return ok, desc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=240.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    ctx = _lib.Context(0)
    t0 = time.time()
    counts = {"pdq": 0, "hamming": 0}
    failures = []
    case = 0
    while time.time() - t0 < args.seconds:
        rng = np.random.default_rng([args.seed, case])
        fn = pdq_case if case % 3 != 2 else hamming_case
        try:
            ok, desc = fn(ctx, rng, case)
        except Exception as e:       # an error return for a valid input is a failure too
            ok, desc = False, {"case": case, "path": "pdq" if fn is pdq_case else "hamming", "error": str(e)[:300]}
        counts[desc["path"]] += 1
        if not ok:
            failures.append(desc)
            print("MISMATCH", json.dumps(desc), flush=True)
            if len(failures) >= 5:
                break
        case += 1
    res = {"seed": args.seed, "seconds": round(time.time() - t0, 1), "cases": case, "by_path": counts,
           "failures": failures, "library": os.path.basename(_lib.SO_PATH)}
    print(json.dumps(res))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)
    ctx.close()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
