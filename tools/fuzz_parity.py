"""Time-bounded randomised parity run of both hot paths through the C ABI against the CPU oracle.

    python tools/fuzz_parity.py [--seconds 240] [--seed 1] [--out gpurun_out/fuzz.json] [--multi]

The fixed cases live in tests/; this tool walks the input space the tests sample: random image
sizes around every kernel's domain borders (landscape fused kernel, float-chain kernel, generic
pipeline, general Box pre-downsample), all three layouts, padded rows / images, image content
from flat to pure noise, and for the grouping path random n, thresholds 0..63, variant counts,
has_hash / low_conf masks and hash populations (uniform, planted clusters, shared prefixes, blocks
of identical hashes); plus the smaller entry points: pHash, u64 grouping, star clustering
(find_groups), hash / dihedral hashes from coefficients (ties, signed zeros, infinities,
denormals), the scanner feeder with mixed sizes, group_max_dist.  Every case is compared bit for bit (hash, quality, coefficient bit
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


def phash_case(ctx, rng, case):
    """rh_phash_batch (+ dihedral set) vs the oracle's restatement of phash.rs:48-83"""
    from rupphash_b200 import phash
    h = int([rng.integers(1, 40), rng.integers(20, 700), rng.integers(700, 1800), 32][int(rng.integers(0, 4))])
    w = int([rng.integers(1, 40), rng.integers(20, 700), rng.integers(700, 1800), 32][int(rng.integers(0, 4))])
    ch = [3, 4, 1][int(rng.integers(0, 3))]
    n = int(rng.integers(1, 5))
    imgs = make_images(rng, n, h, w, ch)
    arr = imgs[..., 0] if ch == 1 else imgs
    desc = dict(case=case, path="phash", h=h, w=w, ch=ch, n=n)
    got, dih = phash.DctPhash(ctx).hash_batch(np.ascontiguousarray(arr), want_dihedral=True)
    ok = True
    for k in range(n):
        want, _ = orc.phash_image(np.ascontiguousarray(arr[k]), layout={3: 0, 4: 1, 1: 2}[ch])
        ok = ok and int(got[k]) == want and [int(x) for x in dih[k]] == orc.phash_dihedral(want)
    return ok, desc


def u64_case(ctx, rng, case):
    """rh_hamming_group_u64 vs the 256-bit oracle on the zero-extended values (same distances)"""
    import ctypes as C
    n = int([rng.integers(1, 70), rng.integers(70, 5000)][int(rng.integers(0, 2))])
    similarity = int(rng.integers(0, 33))
    u = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    k = n // 3
    if k:          # near copies of other values, both sides of the threshold
        src = rng.integers(0, n, size=k)
        dst = rng.integers(0, n, size=k)
        for a, b in zip(src, dst):
            bits = rng.choice(64, size=int(rng.integers(0, min(64, similarity + 6))), replace=False)
            m = np.uint64(0)
            for bpos in bits:
                m |= np.uint64(1) << np.uint64(bpos)
            u[b] = u[a] ^ m
    has_hash = (rng.random(n) < 0.9).astype(np.uint8) if rng.integers(0, 3) == 0 else None
    lc = (rng.random(n) < 0.2).astype(np.uint8) if rng.integers(0, 2) else None
    variants = n_variants = None
    if rng.integers(0, 2):
        variants = rng.integers(0, 1 << 63, size=(n, 8), dtype=np.uint64)
        variants[:, 0] = u
        for a in rng.integers(0, n, size=max(1, n // 10)):
            variants[a, int(rng.integers(1, 8))] = u[int(rng.integers(0, n))] ^ np.uint64(int(rng.integers(0, 8)))
        if rng.integers(0, 2):
            n_variants = rng.integers(0, 9, size=n).astype(np.uint8)
    desc = dict(case=case, path="u64", n=n, similarity=similarity, variants=variants is not None,
                n_variants=n_variants is not None, has_hash=has_hash is not None, low_conf=lc is not None)
    wide = np.zeros((n, 32), np.uint8)
    wide[:, :8] = u.view(np.uint8).reshape(n, 8)
    wide_var = None
    if variants is not None:
        wide_var = np.zeros((n, 8, 32), np.uint8)
        wide_var[:, :, :8] = np.ascontiguousarray(variants).view(np.uint8).reshape(n, 8, 8)
    ref_labels, ref_cnt, _ = orc.group_generic(wide, similarity, has_hash=has_hash, variants=wide_var,
                                               n_variants=n_variants, low_conf=lc, use_mih=False)
    labels = np.empty(n, np.uint32)
    cnt = C.c_uint64()
    ctx.check(_lib.lib().rh_hamming_group_u64(ctx.handle, _lib.ptr(u), _lib.ptr(has_hash), _lib.ptr(variants),
                                              _lib.ptr(n_variants), _lib.ptr(lc), n, similarity, _lib.ptr(labels),
                                              C.byref(cnt)))
    return cnt.value == ref_cnt and np.array_equal(labels, ref_labels), desc


def find_groups_case(ctx, rng, case):
    """rh_find_groups (greedy star clustering, hamminghash.rs:191-271) vs the oracle's MIH version, inside the
    radius where the reference's probing is complete"""
    from rupphash_b200 import hamminghash
    n = int(rng.integers(2, 5000))
    if rng.integers(0, 2):
        max_dist = int(rng.integers(0, 32))
        hashes, _, pop = make_hashes(rng, n, max_dist)
        if pop in ("few_values", "dense_ball"):
            hashes = hashes[:1500]
    else:
        max_dist = int(rng.integers(0, 16))
        hashes = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
        k = n // 3
        hashes[rng.integers(0, n, size=k)] = hashes[rng.integers(0, n, size=k)] ^ np.uint64(int(rng.integers(0, 1 << 10)))
        pop = "u64"
    desc = dict(case=case, path="find_groups", n=int(len(hashes)), max_dist=max_dist, population=pop)
    got = hamminghash.find_groups(hamminghash.MIHIndex.new(hashes), max_dist, ctx)
    want = orc.MIHIndex(hashes).find_groups(max_dist, threads=8)
    return [g[0] for g in got] == [g[0] for g in want] and [sorted(g) for g in got] == [sorted(g) for g in want], desc


def coeffs_case(ctx, rng, case):
    """rh_pdq_hash_from_coeffs / rh_pdq_dihedral_from_coeffs / rh_pdq_from_buffer64 on awkward values"""
    from rupphash_b200 import pdqhash
    n = int(rng.integers(1, 200))
    kind = int(rng.integers(0, 5))
    coeffs = (rng.standard_normal((n, 256)) * float(10.0 ** rng.integers(-3, 6))).astype(np.float32)
    if kind == 1:      # heavy ties around the median
        coeffs = np.round(coeffs / np.float32(coeffs.std() + 1e-9) * 2).astype(np.float32)
    elif kind == 2:    # signed zeros, infinities
        m = rng.random((n, 256))
        coeffs[m < 0.3] = 0.0
        coeffs[(m >= 0.3) & (m < 0.5)] = -0.0
        coeffs[m > 0.98] = np.inf
        coeffs[(m > 0.96) & (m <= 0.98)] = -np.inf
    elif kind == 3:    # denormals
        coeffs = (coeffs * np.float32(1e-42)).astype(np.float32)
    desc = dict(case=case, path="coeffs", n=n, kind=kind)
    h = pdqhash.hash_from_coeffs(coeffs, ctx)
    d = pdqhash.dihedral_from_coeffs(coeffs, ctx)
    ok = all(np.array_equal(h[k], orc.to_hash(coeffs[k])) and np.array_equal(d[k], orc.dihedral(coeffs[k]))
             for k in range(n))
    m = int(rng.integers(1, 12))
    bufs = (rng.random((m, 64, 64), dtype=np.float32) * np.float32(255.0)).astype(np.float32)
    if rng.integers(0, 2):
        bufs = np.round(bufs / 32).astype(np.float32) * 32
    got = pdqhash.from_buffer64(bufs, want_coeffs=True, want_dihedral=True, ctx=ctx)
    for k in range(m):
        c = orc.dct64_to_16(bufs[k])
        ok = (ok and np.array_equal(got["coeffs"][k].view(np.uint32), c.view(np.uint32).reshape(-1))
              and np.float32(orc.quality(bufs[k])) == got["quality"][k] and np.array_equal(got["hash"][k], orc.to_hash(c))
              and np.array_equal(got["dihedral"][k], orc.dihedral(c)))
    return ok, desc


def feeder_case(ctx, rng, case):
    """scanner.hash_files_batched (pinned staging, async submits, mixed sizes in arrival order) vs the oracle"""
    shapes = [pick_shape(rng) for _ in range(int(rng.integers(1, 4)))]
    shapes = [(min(h, 1100), min(w, 1100)) for h, w in shapes]
    ch = [3, 4, 1][int(rng.integers(0, 3))]
    pools = [make_images(rng, int(rng.integers(1, 9)), h, w, ch) for h, w in shapes]
    order = [(s, i) for s, pool in enumerate(pools) for i in range(len(pool))]
    rng.shuffle(order)
    imgs = [pools[s][i] if ch > 1 else pools[s][i][..., 0] for s, i in order]
    batch = int(rng.integers(1, 9))
    workers = int(rng.integers(0, 4))
    desc = dict(case=case, path="feeder", shapes=shapes, ch=ch, images=len(imgs), batch_size=batch, workers=workers)
    if workers:
        res = scanner.hash_files_batched(list(range(len(imgs))), batch_size=batch, ctx=ctx, decode=lambda k: imgs[k],
                                         workers=workers)
    else:
        res = scanner.hash_files_batched(iter(imgs), batch_size=batch, ctx=ctx)
    wants = [orc.pdq_batch(pool, layout={3: 0, 4: 1, 1: 2}[ch], threads=8, want_coeffs=True) for pool in pools]
    ok = len(res) == len(imgs)
    for k, (s, i) in enumerate(order):
        if not ok:
            break
        w = wants[s]
        if not w["valid"][i]:
            ok = res[k] is None
        else:
            ok = (res[k] is not None and np.array_equal(res[k]["hash"], w["hash"][i])
                  and np.float32(res[k]["quality"]) == w["quality"][i] and np.array_equal(res[k]["coeffs"], w["coeffs"][i]))
    return ok, desc


def max_dist_case(ctx, rng, case):
    """rh_group_max_dist vs the reference rule (scanner.rs:2217-2241)"""
    n = int(rng.integers(8, 600))
    hashes, _, _ = make_hashes(rng, n, 31)
    n = len(hashes)
    coeffs = (rng.standard_normal((n, 256)) * 30).astype(np.float32)
    has_hash = (rng.random(n) > 0.1).astype(np.uint8)
    has_hash[0] = 1
    groups = []
    for _ in range(int(rng.integers(1, 40))):
        g = sorted(rng.choice(n, size=int(rng.integers(1, min(n, 12))), replace=False).tolist())
        if any(has_hash[i] for i in g):
            groups.append(g)
    pivots = [next(i for i in g if has_hash[i]) for g in groups]
    desc = dict(case=case, path="max_dist", n=n, groups=len(groups))
    got = scanner.group_max_dist(groups, hashes, pivots, coefficients=coeffs, has_hash=has_hash, ctx=ctx)
    got_plain = scanner.group_max_dist(groups, hashes, pivots, has_hash=has_hash, ctx=ctx)
    ok = True
    for g, members in enumerate(groups):
        variants = orc.dihedral(coeffs[pivots[g]])
        want = max(min(orc.hamming256(v, hashes[i]) for v in variants) for i in members if has_hash[i])
        want_plain = max(orc.hamming256(hashes[pivots[g]], hashes[i]) for i in members if has_hash[i])
        ok = ok and got[g] == want and got_plain[g] == want_plain
    return ok, desc


def multi_group_case(groups, rng, case):
    """rh_hamming_group_multi over all GPUs of the box (one process) vs the oracle; the rh_group is one of
    several created with different flags (NCCL / peer copies, static / stolen tiles)"""
    import torch
    flags, grp = groups[int(rng.integers(0, len(groups)))]
    n = int([rng.integers(1, 70), rng.integers(70, 3000), rng.integers(3000, 12000)][int(rng.integers(0, 3))])
    similarity = int([rng.integers(0, 32), rng.integers(32, 64), 63, 31, 40][int(rng.integers(0, 5))])
    hashes, low_conf, pop = make_hashes(rng, n, similarity)
    if pop in ("few_values", "dense_ball") and n > 4000:
        n = 4000
        hashes, low_conf = hashes[:n], low_conf[:n]
    variants = random_variants(hashes, seed=int(rng.integers(1, 1 << 30))) if rng.integers(0, 2) else None
    n_variants = rng.integers(0, 9, size=n).astype(np.uint8) if variants is not None and rng.integers(0, 2) else None
    has_hash = (rng.random(n) < 0.9).astype(np.uint8) if rng.integers(0, 3) == 0 else None
    lc = low_conf if rng.integers(0, 2) else None
    on_device = bool(rng.integers(0, 2))
    dev = int(rng.integers(0, grp.size))
    desc = dict(case=case, path="multi_group", flags=flags, n=n, similarity=similarity, population=pop,
                variants=variants is not None, n_variants=n_variants is not None, has_hash=has_hash is not None,
                low_conf=lc is not None, device_resident=on_device, device=dev)
    ref_labels, ref_cnt, _ = orc.group_generic(hashes, similarity, has_hash=has_hash, variants=variants,
                                               n_variants=n_variants, low_conf=lc, use_mih=False)
    put = (lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to("cuda:%d" % dev)) if on_device else (lambda a: a)
    labels, cnt = scanner.group_labels_multi(grp, put(hashes), similarity, has_hash=put(has_hash), variants=put(variants),
                                             n_variants=put(n_variants), low_conf=put(lc))
    return cnt == ref_cnt and np.array_equal(labels, ref_labels), desc


def multi_pdq_case(groups, rng, case):
    """rh_pdq_hash_batch_multi: one batch split over the GPUs of the group vs the oracle"""
    _, grp = groups[int(rng.integers(0, len(groups)))]
    h, w = pick_shape(rng)
    h, w = min(h, 1100), min(w, 1100)
    ch = [3, 4, 1][int(rng.integers(0, 3))]
    layout = {3: 0, 4: 1, 1: 2}[ch]
    n = int(rng.integers(1, 24))
    imgs = make_images(rng, n, h, w, ch)
    desc = dict(case=case, path="multi_pdq", h=h, w=w, ch=ch, n=n)
    want = orc.pdq_batch(imgs, layout=layout, threads=16, want_coeffs=True, want_dihedral=True)
    got = {"hash": np.zeros((n, 32), np.uint8), "quality": np.zeros(n, np.float32),
           "coeffs": np.zeros((n, 256), np.float32), "dihedral": np.zeros((n, 8, 32), np.uint8),
           "valid": np.zeros(n, np.uint8)}
    grp.check(_lib.lib().rh_pdq_hash_batch_multi(grp.handle, _lib.ptr(imgs), layout, n, w, h, 0, 0, _lib.ptr(got["hash"]),
                                                 _lib.ptr(got["quality"]), _lib.ptr(got["coeffs"]),
                                                 _lib.ptr(got["dihedral"]), _lib.ptr(got["valid"])))
    ok = (np.array_equal(got["valid"], want["valid"]) and np.array_equal(got["hash"], want["hash"])
          and np.array_equal(got["quality"].view(np.uint32), want["quality"].view(np.uint32))
          and np.array_equal(got["coeffs"].view(np.uint32), want["coeffs"].view(np.uint32))
          and np.array_equal(got["dihedral"], want["dihedral"]))
    return ok, desc


CASES = [pdq_case, pdq_case, hamming_case, pdq_case, phash_case, hamming_case, u64_case, pdq_case, find_groups_case,
         coeffs_case, feeder_case, max_dist_case]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=240.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", default="")
    ap.add_argument("--multi", action="store_true", help="the rh_group entry points over every GPU of the box")
    args = ap.parse_args()
    cases = CASES
    if args.multi:
        ctx = [(f, _lib.Group(flags=f)) for f in (0, _lib.GROUP_NO_NCCL, _lib.GROUP_STEAL_TILES,
                                                   _lib.GROUP_NO_NCCL | _lib.GROUP_STEAL_TILES)]
        cases = [multi_group_case, multi_group_case, multi_pdq_case]
    else:
        ctx = _lib.Context(0)
    t0 = time.time()
    counts = {}
    failures = []
    case = 0
    while time.time() - t0 < args.seconds:
        rng = np.random.default_rng([args.seed, case])
        fn = cases[case % len(cases)]
        try:
            ok, desc = fn(ctx, rng, case)
        except Exception as e:       # an error return for a valid input is a failure too
            ok, desc = False, {"case": case, "path": fn.__name__[:-5], "error": repr(e)[:300]}
        counts[desc["path"]] = counts.get(desc["path"], 0) + 1
        if not ok:
            failures.append(desc)
            print("MISMATCH", json.dumps(desc), flush=True)
            if len(failures) >= 5:
                break
        case += 1
    res = {"seed": args.seed, "seconds": round(time.time() - t0, 1), "cases": case, "by_path": counts,
           "failures": failures, "library": os.path.basename(_lib.SO_PATH)}
    if args.multi:
        res["gpus"] = ctx[0][1].size
    print(json.dumps(res))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)
    if args.multi:
        for _, g in ctx:
            g.close()
    else:
        ctx.close()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
