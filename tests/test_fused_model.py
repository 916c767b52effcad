"""CPU proofs behind the fused PDQ kernel (rupphash_b200/csrc/pdq_fused.cu):
  * the two-term-reciprocal division it uses equals IEEE f32 division on the whole finite input set;
  * the restructured algorithm (integer 2-D box sums + one rounding, real chains only where the
    data is inexact) reproduces the oracle's 64x64 buffer bit for bit."""
import os
import sys
from fractions import Fraction

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

f32 = np.float32


def _rn(v: Fraction) -> np.float32:
    """round a rational to f32, ties to even (exact: goes through the integer significand)"""
    if v == 0:
        return f32(0)
    sign = -1 if v < 0 else 1
    v = abs(v)
    e = v.numerator.bit_length() - v.denominator.bit_length()
    if Fraction(2) ** e > v:
        e -= 1
    scaled = v / Fraction(2) ** (e - 23)          # in [2^23, 2^24)
    n = scaled.numerator // scaled.denominator
    rem = scaled - n
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and n % 2 == 1):
        n += 1
    return f32(sign * float(n) * 2.0 ** (e - 23))


@pytest.mark.parametrize("cnt", [1, 2, 3, 4, 5, 6, 7, 8])
def test_fma_division_equals_ieee_division(cnt):
    """div_exact / recip2 of pdq_fused.cu: yh = RN(1/d), yl = RN(RN(1 - d yh) yh),
    q = fma(f, yh, RN(f yl)) against IEEE division, every step in exact rational arithmetic."""
    for scale, smax in ((8, 8 * 8 * 255), (4, 8 * 4 * 255)):
        d = f32(scale * cnt)
        yh = f32(f32(1) / d)                                                 # __frcp_rn
        r = _rn(Fraction(1) - Fraction(float(d)) * Fraction(float(yh)))      # fma(-d, yh, 1)
        yl = _rn(Fraction(float(r)) * Fraction(float(yh)))
        s = np.arange(0, smax + 1)
        want = (s.astype(f32) / d).astype(f32)
        # vectorised: f yh is exact in f64 (48 bits), t = RN(f yl) has 24; the sum spans < 53 bits
        t = (s.astype(np.float64) * np.float64(yl)).astype(f32)
        q = (s.astype(np.float64) * np.float64(yh) + t.astype(np.float64)).astype(f32)
        assert np.array_equal(q, want), (cnt, scale)
        for k in range(0, smax + 1, 61):                                     # the same through Fractions
            tk = _rn(Fraction(k) * Fraction(float(yl)))
            got = _rn(Fraction(k) * Fraction(float(yh)) + Fraction(float(tk)))
            assert got == want[k], (cnt, scale, k)


@pytest.mark.parametrize("d", [5, 6, 7])
def test_edge_column_quotients_equal_ieee_division(d):
    """div_small of pdq_fused.cu (pass-1 values of the six clipped columns, pdqhash.rs:372-378, :389-395):
    every sum of up to seven u8 pixels over 5, 6 and 7, in exact rational arithmetic."""
    dd = f32(d)
    yh = f32(f32(1) / dd)
    r = _rn(Fraction(1) - Fraction(float(dd)) * Fraction(float(yh)))
    yl = _rn(Fraction(float(r)) * Fraction(float(yh)))
    for k in range(0, 7 * 255 + 1):
        tk = _rn(Fraction(k) * Fraction(float(yl)))
        got = _rn(Fraction(k) * Fraction(float(yh)) + Fraction(float(tk)))
        assert got == f32(f32(k) / dd), (d, k)


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6, 7, 8])
def test_pass1_quotients_of_the_float_kernel(d):
    """pass1_pair of pdq_float.cu: window sums of up to eight u8 pixels over the clipped window size 1..8,
    by the same two-term reciprocal, in exact rational arithmetic."""
    dd = f32(d)
    yh = f32(f32(1) / dd)
    r = _rn(Fraction(1) - Fraction(float(dd)) * Fraction(float(yh)))
    yl = _rn(Fraction(float(r)) * Fraction(float(yh)))
    for k in range(0, 8 * 255 + 1):
        tk = _rn(Fraction(k) * Fraction(float(yl)))
        got = _rn(Fraction(k) * Fraction(float(yh)) + Fraction(float(tk)))
        assert got == f32(f32(k) / dd), (d, k)


def test_restructured_algorithm_is_bit_exact(orc):
    import fused_model
    from rupphash_b200.synth import synth_images
    rng = np.random.default_rng(1)
    # column windows 6, 6, 8 (the BASELINE shapes) and 4, 5, 7 (the other heights the fused kernel accepts)
    for (h, w) in [(384, 512), (341, 512), (512, 512), (193, 512), (288, 512), (448, 512)]:
        for luma in (orc.luma601(synth_images(1, h, w, seed=h)[0]).reshape(h, w),
                     rng.integers(0, 256, size=(h, w), dtype=np.uint8)):
            _, _, ref = orc.pdq_from_luma(luma)
            got = fused_model.fused_buffer64(luma)
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_fp32_fma_divides_luma_by_1000_exactly():
    """pdq_fused.cu luma_px: floor(v / 1000) + 1 from one FP32 FMA on the DP2A accumulator, for every
    v = 299 r + 587 g + 114 b + 500 the kernel can see.  The FMA is evaluated here in exact integer
    arithmetic (mantissa of the f32 constant x F, one round-to-nearest-even at the end)."""
    c = np.float32(0.0005)
    mant, exp = np.frexp(c)                       # c = mant * 2**exp, mant in [0.5, 1)
    m24 = int(np.ldexp(mant, 24))                 # 24-bit integer mantissa
    shift = 24 - int(exp)                         # c = m24 / 2**shift
    assert np.float32(m24 / 2.0 ** shift) == c
    v = np.arange(0, 256001, dtype=np.int64)
    F = 8388608 + 1392 + 1001 + 2 * v             # the float the DP2A pair leaves, as an integer
    assert F.max() < 2 ** 24
    exact = F * m24 + (8384413 << shift)          # (F * c + M2) * 2**shift, exact
    assert (exact >> shift).min() >= 2 ** 23 and (exact >> shift).max() < 2 ** 24   # ulp(result) == 1
    half = 1 << (shift - 1)
    frac = exact & ((1 << shift) - 1)
    assert (frac != half).all()                   # no ties
    rn = (exact + half) >> shift
    assert np.array_equal(rn - 8388608, v // 1000 + 1)
