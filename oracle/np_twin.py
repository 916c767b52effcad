"""Independent numpy restatement of /root/reference/src/pdqhash.rs (second oracle).

TEST INFRASTRUCTURE ONLY.  Written separately from oracle_pdq.c (different language,
different loop structure: every 1-D box chain is advanced in lock-step across all
lines as float32 vectors) so that the two agreeing bit-for-bit is evidence that both
follow the reference.  All arithmetic is np.float32; numpy never fuses mul+add.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def luma601(rgb: np.ndarray) -> np.ndarray:
    """pdqhash.rs:268-284; rgb (h, w, >=3) uint8 -> (h, w) uint8"""
    r = rgb[..., 0].astype(np.uint32)
    g = rgb[..., 1].astype(np.uint32)
    b = rgb[..., 2].astype(np.uint32)
    return ((299 * r + 587 * g + 114 * b + 500) // 1000).astype(np.uint8)


def target_dimensions(w: int, h: int, max_dim: int = 512):
    """pdqhash.rs:224-235"""
    if w == 0 or h == 0:
        return max(w, 1), max(h, 1)
    if w > h:
        return max_dim, max(h * max_dim // w, 1)
    return max(w * max_dim // h, 1), max_dim


def downsample_2x(luma: np.ndarray) -> np.ndarray:
    """fast_image_resize Box at an exact 2x ratio: horizontal (a+b+1)>>1 into u8, then vertical."""
    a = luma.astype(np.uint16)
    hz = (a[:, 0::2] + a[:, 1::2] + 1) >> 1
    return ((hz[0::2, :] + hz[1::2, :] + 1) >> 1).astype(np.uint8)


def box_weight_matrix(in_size: int, out_size: int):
    """fast_image_resize 6.1.0 `Convolution(FilterType::Box)` coefficients for one axis, written independently of
    oracle_pdq.c's sparse (xmin, count, taps) loops: a DENSE (out, in) matrix built by broadcasting.
    Pillow-derived rule: scale = in/out (>= 1 when shrinking), support = scale / 2, an input pixel x belongs to
    output o when its centre x + 0.5 lies in (centre - support, centre + support] with centre = (o + 0.5) scale,
    restricted to the integer span [floor(centre - support + 0.5), floor(centre + support + 0.5)); equal weights,
    normalised per output; fixed point at the largest precision that keeps every coefficient below 2^15.
    The crate is not in the reference tree: parity with it is unpinned.  -> (int matrix, precision)"""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 0.5 * fscale
    o = np.arange(out_size, dtype=np.float64)[:, None]
    x = np.arange(in_size, dtype=np.float64)[None, :]
    centre = (o + 0.5) * scale
    lo = np.maximum(np.floor(centre - support + 0.5), 0.0)
    hi = np.minimum(np.floor(centre + support + 0.5), float(in_size))
    t = (x - centre + 0.5) / fscale
    member = (x >= lo) & (x < hi) & (t > -0.5) & (t <= 0.5)
    wgt = member.astype(np.float64)
    tot = wgt.sum(axis=1, keepdims=True)
    wgt = np.divide(wgt, tot, out=np.zeros_like(wgt), where=tot != 0)
    biggest = float(wgt.max())
    precision = 0
    while precision < 22 and int(0.5 + biggest * float(1 << (precision + 1))) < (1 << 15):
        precision += 1
    scaled = wgt * float(1 << precision)
    k = np.where(scaled < 0, np.ceil(scaled - 0.5), np.floor(scaled + 0.5)).astype(np.int64)
    return k, precision


def resize_box_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """pdqhash.rs:203-220 (resize_luma_fast): horizontal pass into a u8 plane, then vertical; each pass is
    (sum(px * k) + 2^(p-1)) >> p clipped to [0, 255]."""
    sh, sw = src.shape
    kx, px = box_weight_matrix(sw, dw)
    ky, py = box_weight_matrix(sh, dh)
    a = src.astype(np.int64)
    hz = np.clip((a @ kx.T + (1 << (px - 1))) >> px, 0, 255)            # (sh, dw)
    out = np.clip((ky @ hz + (1 << (py - 1))) >> py, 0, 255)            # (dh, dw)
    return out.astype(np.uint8)


def box_lines(x: np.ndarray, win: int) -> np.ndarray:
    """pdqhash.rs:341-396 applied along axis 1 of a (lines, len) float32 array."""
    lines, n = x.shape
    win = min(max(win, 1), max(n, 1))
    half = (win + 2) // 2
    p1, p2, p3, p4 = half - 1, win - half + 1, max(n - win, 0), half - 1
    out = np.zeros_like(x)
    s = np.zeros(lines, F)
    cw = F(0.0)
    li = ri = oi = 0
    for _ in range(p1):
        s = s + x[:, ri]
        cw = F(cw + F(1.0))
        ri += 1
    for _ in range(p2):
        s = s + x[:, ri]
        cw = F(cw + F(1.0))
        out[:, oi] = s / cw
        ri += 1
        oi += 1
    for _ in range(p3):
        s = s + x[:, ri]
        s = s - x[:, li]
        out[:, oi] = s / cw
        li += 1
        ri += 1
        oi += 1
    for _ in range(p4):
        s = s - x[:, li]
        cw = F(cw - F(1.0))
        out[:, oi] = s / cw
        li += 1
        oi += 1
    return out


def jarosz(plane: np.ndarray, w_rows: int, w_cols: int, nreps: int = 2) -> np.ndarray:
    """pdqhash.rs:410-426"""
    buf = plane.astype(F)
    for _ in range(nreps):
        tmp = box_lines(buf, w_rows)
        buf = np.ascontiguousarray(box_lines(np.ascontiguousarray(tmp.T), w_cols).T)
    return buf


def decimate64(plane: np.ndarray) -> np.ndarray:
    """pdqhash.rs:428-443"""
    rows, cols = plane.shape
    ri = ((np.arange(64) * 2 + 1) * rows) // 128
    ci = ((np.arange(64) * 2 + 1) * cols) // 128
    return plane[np.ix_(ri, ci)].astype(F)


def quality(buf: np.ndarray) -> float:
    """pdqhash.rs:445-460.  Every term is a small integer, so the f32 sum is exact."""
    buf = buf.astype(F)
    v = np.trunc(np.abs(((buf[:-1, :] - buf[1:, :]) * F(100.0)) / F(255.0)))
    h = np.trunc(np.abs(((buf[:, :-1] - buf[:, 1:]) * F(100.0)) / F(255.0)))
    s = F(0.0)
    for t in np.concatenate([v.ravel(), h.ravel()]):
        s = F(s + t)
    q = F(s / F(90.0))
    return float(min(q, F(1.0)))


def dct_matrix(cosf=None) -> np.ndarray:
    """pdqhash.rs:287-304.  cosf: f32->f32 cosine; default rounds the f64 cosine."""
    if cosf is None:
        cosf = lambda a: F(np.cos(np.float64(a)))
    pi = F(np.pi)
    inv = F(1.0) / np.sqrt(F(64.0))
    norm = F(inv * np.sqrt(F(2.0)))
    d = np.zeros((16, 64), F)
    for i in range(16):
        freq = F(i + 1)
        for j in range(64):
            angle = F(F(F(pi * freq) * F(F(2.0) * F(j) + F(1.0))) / F(128.0))
            d[i, j] = F(norm * cosf(angle))
    return d


def dct64_to_16(buf: np.ndarray, d: np.ndarray) -> np.ndarray:
    """pdqhash.rs:306-336, k-ascending accumulation from 0.0"""
    buf = buf.astype(F)
    inter = np.zeros((16, 64), F)
    for k in range(64):
        inter = inter + d[:, k:k + 1] * buf[k:k + 1, :]
    out = np.zeros((16, 16), F)
    for k in range(64):
        out = out + inter[:, k:k + 1] * d[:, k][None, :]
    return out.reshape(256)


def _signed(c: np.ndarray, neg_rows: bool, neg_cols: bool) -> np.ndarray:
    """pdqhash.rs:127-137: negate where the DCT *frequency* (index+1) is odd"""
    m = c.reshape(16, 16).copy()
    fr = ((np.arange(16) + 1) % 2 == 1) & neg_rows
    fc = ((np.arange(16) + 1) % 2 == 1) & neg_cols
    flip = fr[:, None] ^ fc[None, :]
    m[flip] = -m[flip]
    return m


def _bits(c, neg_rows, neg_cols) -> np.ndarray:
    m = _signed(c, neg_rows, neg_cols)
    median = np.sort(m.ravel(), kind="stable")[127]  # pdqhash.rs:122-123
    return m > median


def _pack(bits: np.ndarray) -> np.ndarray:
    """pdqhash.rs:155-162: coefficient n -> bit n%8 of byte 31 - n//8"""
    flat = bits.reshape(256)
    out = np.zeros(32, np.uint8)
    for n in np.nonzero(flat)[0]:
        out[31 - n // 8] |= np.uint8(1 << (n % 8))
    return out


def to_hash(c: np.ndarray) -> np.ndarray:
    return _pack(_bits(c, False, False))


def dihedral(c: np.ndarray) -> np.ndarray:
    """pdqhash.rs:71-87"""
    idb, nc, nr, nb = _bits(c, False, False), _bits(c, False, True), _bits(c, True, False), _bits(c, True, True)
    return np.stack([_pack(idb), _pack(nr.T), _pack(nb), _pack(nc.T), _pack(nc), _pack(nr), _pack(idb.T), _pack(nb.T)])


def pdq_from_luma(luma: np.ndarray, d: np.ndarray):
    """pdqhash.rs:238-262 -> (coeffs, quality, buf64)"""
    rows, cols = luma.shape
    plane = jarosz(luma.astype(F), -(-cols // 64), -(-rows // 64))
    buf = decimate64(plane)
    return dct64_to_16(buf, d), quality(buf), buf


def pdq_features(img: np.ndarray, d: np.ndarray):
    """pdqhash.rs:166-196 for RGB8 / RGBA8 / Luma8 arrays"""
    h, w = img.shape[:2]
    if w < 5 or h < 5:
        return None
    luma = img if img.ndim == 2 else luma601(img)
    if w > 512 or h > 512:
        nw, nh = target_dimensions(w, h)
        luma = downsample_2x(luma) if (nw * 2, nh * 2) == (w, h) else resize_box_u8(luma, nw, nh)
    return pdq_from_luma(luma, d)
