"""Pins oracle/oracle_pdq.c with every portable unit test of /root/reference/src/pdqhash.rs
(tests at :462-648) plus an independent numpy twin.  CPU only."""
import numpy as np
import pytest

from oracle import np_twin


def lcg_features(seed):
    """pdqhash.rs:537-545"""
    state = seed & 0xFFFFFFFF
    c = np.zeros(256, np.float32)
    for i in range(256):
        state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
        c[i] = np.float32(np.float32(state >> 8) / np.float32(65536.0)) - np.float32(128.0)
    return c


def lcg_buffer(seed):
    """pdqhash.rs:606-614"""
    state = seed & 0xFFFFFFFF
    buf = np.zeros((64, 64), np.float32)
    for r in range(64):
        for c in range(64):
            state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
            buf[r, c] = np.float32((state >> 16) & 0xFF)
    return buf


# naive ground truth, pdqhash.rs:470-535, written independently of the oracle
def naive_to_hash(c):
    median = np.sort(c)[(256 - 1) // 2]
    h = np.zeros(32, np.uint8)
    for i in range(32):
        byte = 0
        for j in range(8):
            if c[i * 8 + j] > median:
                byte |= 1 << j
        h[32 - i - 1] = byte
    return h


def naive_transpose(c):
    return np.ascontiguousarray(c.reshape(16, 16).T).reshape(256)


def naive_flip_x(c):
    m = c.reshape(16, 16).copy()
    for col in range(16):
        if (col + 1) % 2 != 0:
            m[:, col] = -m[:, col]
    return m.reshape(256)


def naive_flip_y(c):
    m = c.reshape(16, 16).copy()
    for r in range(16):
        if (r + 1) % 2 != 0:
            m[r, :] = -m[r, :]
    return m.reshape(256)


def naive_dihedral(c):
    T, X, Y = naive_transpose, naive_flip_x, naive_flip_y
    return np.stack([
        naive_to_hash(c), naive_to_hash(X(T(c))), naive_to_hash(Y(X(c))), naive_to_hash(Y(T(c))),
        naive_to_hash(X(c)), naive_to_hash(Y(c)), naive_to_hash(T(c)), naive_to_hash(Y(X(T(c)))),
    ])


@pytest.mark.parametrize("seed", [1, 42, 0x12345678, 0xDEADBEEF])
def test_fast_dihedral_matches_naive(orc, seed):
    """pdqhash.rs:547-558"""
    c = lcg_features(seed)
    assert np.array_equal(orc.to_hash(c), naive_to_hash(c))
    assert np.array_equal(orc.dihedral(c), naive_dihedral(c))
    # the numpy twin agrees too
    assert np.array_equal(np_twin.to_hash(c), naive_to_hash(c))
    assert np.array_equal(np_twin.dihedral(c), naive_dihedral(c))


def test_dihedral_set_is_the_full_group(orc):
    """pdqhash.rs:560-570"""
    hs = orc.dihedral(lcg_features(7))
    for i in range(8):
        for j in range(i + 1, 8):
            assert not np.array_equal(hs[i], hs[j])


def transform(buf, variant):
    """pdqhash.rs:587-604"""
    n = 64
    out = np.zeros_like(buf)
    for x in range(n):
        for y in range(n):
            out[x, y] = [
                lambda: buf[x, y], lambda: buf[n - 1 - y, x], lambda: buf[n - 1 - x, n - 1 - y],
                lambda: buf[y, n - 1 - x], lambda: buf[x, n - 1 - y], lambda: buf[n - 1 - x, y],
                lambda: buf[y, x], lambda: buf[n - 1 - y, n - 1 - x],
            ][variant]()
    return out


@pytest.mark.parametrize("seed", [1, 42, 0xDEADBEEF])
def test_dihedral_hashes_match_physically_transformed_buffer(orc, seed):
    """pdqhash.rs:582-628: distance 0 for all 8 variants through the real DCT"""
    buf = lcg_buffer(seed)
    predicted = orc.dihedral(orc.dct64_to_16(buf))
    for variant in range(8):
        actual = orc.to_hash(orc.dct64_to_16(transform(buf, variant)))
        assert orc.hamming256(actual, predicted[variant]) == 0, f"variant {variant}"


def test_quality_metric_scaling(orc):
    """pdqhash.rs:630-639"""
    assert orc.quality(np.full((64, 64), 128.0, np.float32)) == 0.0
    buf = np.array([[0.0, 10.0], [0.0, 10.0]], np.float32)
    assert abs(orc.quality(buf) - 6.0 / 90.0) < 1e-6


def test_target_dimensions_never_collapse_to_zero(orc):
    """pdqhash.rs:641-647"""
    assert orc.target_dimensions(4000, 5, 512) == (512, 1)
    assert orc.target_dimensions(5, 4000, 512) == (1, 512)
    assert orc.target_dimensions(1024, 1024, 512) == (512, 512)
    assert orc.target_dimensions(1024, 512, 512) == (512, 256)
    assert np_twin.target_dimensions(4000, 5) == (512, 1)


def test_min_hashable_dim(orc):
    """pdqhash.rs:167-169"""
    assert orc.pdq_features(np.zeros((4, 100, 3), np.uint8)) is None
    assert orc.pdq_features(np.zeros((100, 4, 3), np.uint8)) is None
    assert orc.pdq_features(np.zeros((5, 5, 3), np.uint8)) is not None


def test_box_one_d_window_coverage(orc):
    """output o covers in[o-(win-half) .. o+half-1] clipped (true mean within f32 error)"""
    rng = np.random.default_rng(3)
    for n, win in [(512, 8), (384, 6), (64, 1), (7, 8), (5, 3), (100, 2), (33, 5)]:
        x = rng.integers(0, 256, n).astype(np.float32)
        got = orc.box_one_d(x, win)
        w = min(max(win, 1), n)
        half = (w + 2) // 2
        for o in range(n):
            lo, hi = max(0, o - (w - half)), min(n - 1, o + half - 1)
            assert abs(got[o] - x[lo:hi + 1].astype(np.float64).mean()) < 1e-3, (n, win, o)
        assert np.array_equal(got, np_twin.box_lines(x[None, :], win)[0])


def test_luma601_exhaustive_sample(orc):
    rng = np.random.default_rng(5)
    px = rng.integers(0, 256, (4096, 3), dtype=np.uint8)
    px[:4] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 0, 255]]
    p32 = px.astype(np.int64)
    want = ((299 * p32[:, 0] + 587 * p32[:, 1] + 114 * p32[:, 2] + 500) // 1000).astype(np.uint8)
    assert np.array_equal(orc.luma601(px), want)
    rgba = np.concatenate([px, rng.integers(0, 256, (4096, 1), dtype=np.uint8)], axis=1)
    assert np.array_equal(orc.luma601(rgba, orc.LAYOUT_RGBA8), want)


def test_resize_2x_is_two_stage_rounded_average(orc):
    rng = np.random.default_rng(11)
    src = rng.integers(0, 256, (768, 1024), dtype=np.uint8)
    assert np.array_equal(orc.resize_box_u8(src, 512, 384), np_twin.downsample_2x(src))
    src = rng.integers(0, 256, (64, 96), dtype=np.uint8)
    assert np.array_equal(orc.resize_box_u8(src, 48, 32), np_twin.downsample_2x(src))


def _glibc_cosf():
    import ctypes

    libm = ctypes.CDLL("libm.so.6")
    libm.cosf.argtypes = [ctypes.c_float]
    libm.cosf.restype = ctypes.c_float
    return lambda a: np.float32(libm.cosf(float(a)))


def test_dct_matrix_follows_platform_cosf(orc):
    """pdqhash.rs:287-304 uses f32::cos = the platform libm cosf.  glibc's cosf is NOT correctly
    rounded for 14 of the 1024 angles (1 ulp off), so the table is defined as "glibc cosf" (stable
    since glibc 2.28) and the CUDA library embeds exactly these values (tools/gen_dct_table.py)."""
    d_c = orc.dct_matrix()
    assert np.array_equal(d_c, np_twin.dct_matrix(_glibc_cosf()))
    d_cr = np_twin.dct_matrix()  # correctly rounded cosine
    ulp = np.abs(d_c.view(np.int32).astype(np.int64) - d_cr.view(np.int32).astype(np.int64))
    assert ulp.max() <= 2 and (ulp > 0).sum() <= 32


@pytest.mark.parametrize("shape", [(384, 512), (512, 512), (768, 1024), (100, 37), (5, 5), (64, 64), (341, 512), (65, 129)])
def test_c_oracle_matches_numpy_twin(orc, shape):
    """two independent restatements agree bit-for-bit"""
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    d = np_twin.dct_matrix(_glibc_cosf())
    for trial in range(3):
        h, w = shape
        if trial == 0:
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        else:
            yy, xx = np.mgrid[0:h, 0:w]
            base = 128 + 60 * np.sin(xx / (7.0 + trial)) + 50 * np.cos(yy / (5.0 + 3 * trial))
            img = np.clip(base[..., None] + rng.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)
        c1, q1, b1 = orc.pdq_features(img)
        c2, q2, b2 = np_twin.pdq_features(img, d)
        assert np.array_equal(b1, b2)
        assert np.array_equal(c1.view(np.uint32), c2.view(np.uint32))
        assert q1 == q2
        assert np.array_equal(orc.to_hash(c1), np_twin.to_hash(c2))


def test_luma_input_is_borrowed(orc):
    """pdqhash.rs:172-175: Luma8 input skips the conversion"""
    rng = np.random.default_rng(8)
    luma = rng.integers(0, 256, (200, 300), dtype=np.uint8)
    c1, q1, _ = orc.pdq_features(luma, orc.LAYOUT_LUMA8)
    c2, q2, _ = orc.pdq_from_luma(luma)
    assert np.array_equal(c1, c2) and q1 == q2


def test_quality_100(orc):
    """scanner.rs:1416-1418"""
    assert orc.quality_100(0.0) == 0 and orc.quality_100(1.0) == 100
    assert orc.quality_100(0.495) == 50 and orc.quality_100(0.494) == 49


def test_batch_mt_matches_single(orc):
    from rupphash_b200.synth import synth_images

    imgs = synth_images(6, 96, 128, seed=4)
    out = orc.pdq_batch(imgs, threads=3, want_coeffs=True, want_dihedral=True)
    for i in range(6):
        c, q, _ = orc.pdq_features(imgs[i])
        assert np.array_equal(out["coeffs"][i], c) and out["quality"][i] == np.float32(q)
        assert np.array_equal(out["hash"][i], orc.to_hash(c))
        assert np.array_equal(out["dihedral"][i], orc.dihedral(c))
        assert out["valid"][i] == 1


@pytest.mark.parametrize("shape,target", [((854, 1280), (512, 341)), ((720, 1080), (512, 341)), ((1280, 854), (341, 512)),
                                          ((513, 513), (512, 512)), ((600, 2000), (512, 153)), ((4000, 5), (1, 512)),
                                          ((1000, 700), (358, 512)), ((768, 1024), (512, 384)), ((97, 53), (31, 57))])
def test_box_resize_two_independent_restatements_agree(orc, shape, target):
    """H5 (fast_image_resize Box convolution): the C oracle's sparse tap loops and the numpy twin's dense weight
    matrix are written independently (the device's coefficient code mirrors the C oracle's, so the
    comparison target has to be a second statement).  Neither is pinned against the real crate."""
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    src = rng.integers(0, 256, size=shape, dtype=np.uint8)
    dw, dh = target
    assert np.array_equal(orc.resize_box_u8(src, dw, dh), np_twin.resize_box_u8(src, dw, dh))
