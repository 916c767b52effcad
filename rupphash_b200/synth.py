"""Seeded synthetic workloads for tests and bench.py.

Pure numpy helpers; nothing here computes a hash or a distance.
"""
from __future__ import annotations

import numpy as np


def planted_hashes(n: int, seed: int = 0xB200, n_clusters: int | None = None, threshold: int = 31,
                   identical_block: int | None = None):
    """n x 32 uint8 PDQ-like hashes: uniform random bits plus planted structure.

    * clusters of 2..8 members at distances 0..threshold+9 around a random centre
      (both sides of the threshold are hit),
    * chains A-B-C with d(A,B), d(B,C) <= threshold < d(A,C) (transitivity),
    * one block of identical hashes flagged low-confidence and one unflagged,
    all shuffled into random positions.  Returns (hashes, low_conf).
    """
    rng = np.random.default_rng(seed)
    hashes = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    low_conf = np.zeros(n, np.uint8)
    if n < 64:
        return hashes, low_conf
    if n_clusters is None:
        n_clusters = max(4, n // 100)
    if identical_block is None:
        identical_block = min(1000, max(2, n // 50))
    perm = rng.permutation(n)
    cursor = 0

    def take(k):
        nonlocal cursor
        idx = perm[cursor:cursor + k]
        cursor += k
        return idx

    def flip(h, nbits):
        out = h.copy()
        if nbits:
            pos = rng.choice(256, size=nbits, replace=False)
            for p in pos:
                out[p >> 3] ^= np.uint8(1 << (p & 7))
        return out

    budget = n // 2
    for _ in range(n_clusters):
        k = int(rng.integers(2, 9))
        if cursor + k > budget:
            break
        idx = take(k)
        centre = hashes[idx[0]].copy()
        for m in idx[1:]:
            hashes[m] = flip(centre, int(rng.integers(0, threshold + 10)))
    for _ in range(max(2, n_clusters // 10)):  # chains
        if cursor + 3 > budget:
            break
        a, b, c = take(3)
        base = hashes[a].copy()
        pos = rng.choice(256, size=2 * threshold, replace=False) if threshold else np.empty(0, int)
        hb = base.copy()
        for p in pos[:threshold]:
            hb[p >> 3] ^= np.uint8(1 << (p & 7))
        hc = hb.copy()
        for p in pos[threshold:]:
            hc[p >> 3] ^= np.uint8(1 << (p & 7))
        hashes[b], hashes[c] = hb, hc
    for flagged in (True, False):  # identical blocks
        k = min(identical_block, max(0, budget - cursor))
        if k < 2:
            break
        idx = take(k)
        hashes[idx] = rng.integers(0, 256, size=32, dtype=np.uint8)
        if flagged:
            low_conf[idx] = 1
    # a few low-confidence singles sitting next to ordinary hashes
    k = min(max(2, n // 200), max(0, budget - cursor) // 2)
    for _ in range(k):
        a, b = take(2)
        hashes[b] = flip(hashes[a], int(rng.integers(1, max(2, threshold))))
        low_conf[b] = 1
    return hashes, low_conf


def random_variants(hashes: np.ndarray, seed: int = 7, near: int = 12):
    """n x 8 x 32 query variants: slot 0 is the hash itself, the rest are random except
    that some variant slots are made to sit `near` bits from some *other* file's hash so
    that variant-only edges exist."""
    rng = np.random.default_rng(seed)
    n = hashes.shape[0]
    var = rng.integers(0, 256, size=(n, 8, 32), dtype=np.uint8)
    var[:, 0, :] = hashes
    k = max(1, n // 20)
    src = rng.integers(0, n, size=k)
    dst = rng.integers(0, n, size=k)
    slot = rng.integers(1, 8, size=k)
    for s, d, v in zip(src, dst, slot):
        h = hashes[d].copy()
        pos = rng.choice(256, size=int(rng.integers(0, near + 1)), replace=False)
        for p in pos:
            h[p >> 3] ^= np.uint8(1 << (p & 7))
        var[s, v] = h
    return var


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synth_images(n: int, h: int, w: int, seed: int = 0xB200, channels: int = 3) -> np.ndarray:
    """(n, h, w, channels) uint8 images: blocky low-frequency field + pixel noise."""
    out = np.empty((n, h, w, channels), np.uint8)
    for i in range(n):
        rng = np.random.default_rng(int(_splitmix64(np.array([seed + i], np.uint64))[0] & np.uint64(0x7FFFFFFF)))
        gh, gw = 24, 32
        field = rng.normal(0.0, 50.0, size=(gh, gw, 1)).astype(np.float32)
        ys = (np.arange(h) * gh // h)[:, None]
        xs = (np.arange(w) * gw // w)[None, :]
        base = field[ys, xs, :]
        noise = rng.normal(0.0, 20.0, size=(h, w, channels)).astype(np.float32)
        offs = rng.normal(0.0, 10.0, size=(1, 1, channels)).astype(np.float32)
        out[i] = np.clip(base + noise + offs + 128.0, 0, 255).astype(np.uint8)
    return out
