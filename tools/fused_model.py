#!/usr/bin/env python
"""numpy model of the restructured ("fused") PDQ front end, checked bit-for-bit against the oracle.

The fused CUDA kernel does not run four float passes.  It relies on three facts about
jarosz_filter_float (pdqhash.rs:410-426) when the row window is 8 (plane width 449..512):
  1. pass 1 (rows, window 8) of u8 luma is exact except in columns 0,1,2 and W-4..W-2 (divisors
     5,6,7): P1 = H/8 with H an integer <= 2040;
  2. pass 2 (columns) of those exact values has exact running sums, so away from the six edge
     columns P2[r][c] = RN(S2d / (8 * cnt_r)) with S2d the integer 2-D box sum -- one rounding;
  3. passes 3 and 4 are genuinely sequential float chains and are run as written, but only the
     64 decimated columns of pass 3 feed pass 4.
The six edge columns get the real sequential column chain.  This script verifies 1-3 and the
FMA-based division first used on the device (q = f*y; r = fma(-d,q,f); q' = fma(r,y,q)); the kernel
now uses the cheaper two-term reciprocal that tests/test_fused_model.py proves equal as well.
"""
import os
import sys
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

f32 = np.float32


def box_window(o, win, n):
    """clipped window [lo, hi] of output o (pdqhash.rs:341-396)"""
    half = (win + 2) // 2
    return max(0, o - (win - half)), min(n - 1, o + half - 1)


def chain_1d(vec, win):
    """box_one_d_float restated with numpy float32 scalars"""
    n = len(vec)
    win = max(1, min(win, max(n, 1)))
    half = (win + 2) // 2
    out = np.zeros(n, f32)
    s = f32(0)
    cw = f32(0)
    li = ri = oi = 0
    for _ in range(half - 1):
        s = f32(s + vec[ri]); cw = f32(cw + 1); ri += 1
    for _ in range(win - half + 1):
        s = f32(s + vec[ri]); cw = f32(cw + 1); out[oi] = f32(s / cw); ri += 1; oi += 1
    for _ in range(max(0, n - win)):
        s = f32(s + vec[ri]); s = f32(s - vec[li]); out[oi] = f32(s / cw); li += 1; ri += 1; oi += 1
    for _ in range(half - 1):
        s = f32(s - vec[li]); cw = f32(cw - 1); out[oi] = f32(s / cw); li += 1; oi += 1
    return out


def fma32(a, b, c):
    """exact fused multiply-add rounded once to f32"""
    v = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
    return f32(float(v)) if True else None


def round_fraction_to_f32(v: Fraction):
    # float(v) rounds correctly to f64; f64 -> f32 double rounding can differ only at exact
    # halfway cases of f32 that are not representable in f64 -- impossible for these magnitudes
    return f32(float(v))


def markstein_div(f, d):
    y = f32(f32(1) / f32(d))
    q = f32(f32(f) * y)
    r = round_fraction_to_f32(-Fraction(float(d)) * Fraction(float(q)) + Fraction(float(f)))
    return round_fraction_to_f32(Fraction(float(r)) * Fraction(float(y)) + Fraction(float(q)))


def check_division():
    bad = 0
    for cnt in range(1, 9):
        for scale in (8, 4):
            d = f32(scale * cnt)
            s = np.arange(0, 8 * 8 * 255 + 1, dtype=np.float32)
            want = (s / d).astype(f32)
            got = np.array([markstein_div(x, d) for x in s[::7]], f32)
            bad += int((got != want[::7]).sum())
    print("markstein division mismatches:", bad)
    return bad == 0


def fused_buffer64(luma):
    """64x64 decimated buffer computed the way the fused kernel does it."""
    H, W = luma.shape
    assert 449 <= W <= 512 and H >= 5
    wr, wc = (W + 63) // 64, (H + 63) // 64
    assert wr == 8
    L = luma.astype(np.int64)
    # horizontal clipped 8-sums and their counts
    cs = np.concatenate([np.zeros((H, 1), np.int64), np.cumsum(L, axis=1)], axis=1)
    lo = np.array([box_window(c, wr, W)[0] for c in range(W)])
    hi = np.array([box_window(c, wr, W)[1] for c in range(W)])
    Hs = cs[:, hi + 1] - cs[:, lo]
    hcnt = hi - lo + 1
    # vertical clipped sums of Hs
    rs = np.concatenate([np.zeros((1, W), np.int64), np.cumsum(Hs, axis=0)], axis=0)
    rlo = np.array([box_window(r, wc, H)[0] for r in range(H)])
    rhi = np.array([box_window(r, wc, H)[1] for r in range(H)])
    S2d = rs[rhi + 1] - rs[rlo]
    vcnt = (rhi - rlo + 1)
    P2 = np.zeros((H, W), f32)
    exact_cols = [c for c in range(W) if hcnt[c] in (8, 4)]
    for c in exact_cols:
        P2[:, c] = (S2d[:, c].astype(f32) / (f32(hcnt[c]) * vcnt.astype(f32))).astype(f32)
    for c in range(W):
        if c in exact_cols:
            continue
        p1 = (Hs[:, c].astype(f32) / f32(hcnt[c])).astype(f32)   # rounded quotients
        P2[:, c] = chain_1d(p1, wc)                               # real sequential chain
    # pass 3: sequential row chains, keep decimated columns
    cols = [((2 * j + 1) * W) // 128 for j in range(64)]
    rows = [((2 * i + 1) * H) // 128 for i in range(64)]
    P3 = np.zeros((H, 64), f32)
    for r in range(H):
        P3[r] = chain_1d(P2[r], wr)[cols]
    # pass 4 on the 64 kept columns, keep decimated rows
    out = np.zeros((64, 64), f32)
    for j in range(64):
        out[:, j] = chain_1d(P3[:, j], wc)[rows]
    return out


def main():
    ok = check_division()
    rng = np.random.default_rng(0)
    from rupphash_b200.synth import synth_images
    cases = []
    for (h, w) in [(384, 512), (512, 512), (341, 512), (100, 449), (65, 480), (5, 512), (64, 500), (200, 511)]:
        img = synth_images(1, h, w, seed=h + w)[0]
        cases.append(oracle.luma601(img).reshape(h, w))
        cases.append(rng.integers(0, 256, size=(h, w), dtype=np.uint8))
    for luma in cases:
        _, _, ref = oracle.pdq_from_luma(luma)
        got = fused_buffer64(luma)
        same = np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        print(luma.shape, "bit-exact" if same else f"MISMATCH max abs {np.abs(got - ref).max()}")
        ok &= same
    print("ALL OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
