"""Cache wire formats (db.rs:47, :224-231, :678-702, :734-761, :1200-1231).  The reference has
no test vectors for them; the known answers below are written out by hand from the format definition
(version byte, postcard varint + little-endian f32)."""
import struct

import numpy as np
import pytest

from rupphash_b200 import cachefmt


def test_pdqhash_value_known_answer():
    h = np.arange(32, dtype=np.uint8)
    v = cachefmt.encode_pdqhash(h)
    assert v == bytes([2]) + bytes(range(32))
    assert np.array_equal(cachefmt.decode_pdqhash(v), h)
    # another algorithm version or a wrong length is a miss, not an error (db.rs:686-694)
    assert cachefmt.decode_pdqhash(bytes([1]) + bytes(32)) is None
    assert cachefmt.decode_pdqhash(bytes([2]) + bytes(31)) is None
    assert cachefmt.decode_pdqhash(b"") is None
    with pytest.raises(ValueError):
        cachefmt.encode_pdqhash(np.zeros(31, np.uint8))


def test_coefficients_value_known_answer():
    c = np.zeros(256, np.float32)
    c[0], c[1], c[255] = 1.0, -2.5, np.float32(3.1415927)
    v = cachefmt.encode_coefficients(c)
    # version, varint(256) = 0x80 0x02, then 256 little-endian f32
    assert v[:3] == bytes([2, 0x80, 0x02]) and len(v) == 3 + 1024
    assert v[3:7] == struct.pack("<f", 1.0) and v[7:11] == struct.pack("<f", -2.5)
    assert v[-4:] == struct.pack("<f", np.float32(3.1415927))
    back = cachefmt.decode_coefficients(v)
    assert back.dtype == np.float32 and np.array_equal(back.view(np.uint32), c.view(np.uint32))
    # short vectors use a one-byte varint
    assert cachefmt.encode_coefficients(np.ones(3, np.float32))[:2] == bytes([2, 3])


def test_coefficients_version_and_corruption():
    v = cachefmt.encode_coefficients(np.ones(256, np.float32))
    assert cachefmt.decode_coefficients(bytes([1]) + v[1:]) is None      # older pipeline: absent (db.rs:752)
    assert cachefmt.decode_coefficients(b"") is None
    with pytest.raises(cachefmt.Corrupted):
        cachefmt.decode_coefficients(v[:-1])                              # truncated payload (db.rs:747)
    with pytest.raises(cachefmt.Corrupted):
        cachefmt.decode_coefficients(bytes([2, 0x80]))                    # truncated varint


def test_varint_boundaries():
    for n, enc in ((0, b"\x00"), (127, b"\x7f"), (128, b"\x80\x01"), (256, b"\x80\x02"), (16384, b"\x80\x80\x01")):
        assert cachefmt._varint(n) == enc
        assert cachefmt._read_varint(enc, 0) == (n, len(enc))


def test_quality_tag():
    assert cachefmt.quality_tag(0.0) == (0xF007, 0)
    assert cachefmt.quality_tag(0.495) == (0xF007, 50)     # (q * 100).round(): half away from zero
    assert cachefmt.quality_tag(1.0) == (0xF007, 100)
    assert cachefmt.quality_tag(7.0) == (0xF007, 100)      # clamp (scanner.rs:1417)


def test_load_cached_mixes_versions_and_gaps():
    h = np.random.default_rng(1).integers(0, 256, size=(4, 32), dtype=np.uint8)
    c = np.random.default_rng(2).normal(size=(4, 256)).astype(np.float32)
    hv = [cachefmt.encode_pdqhash(h[0]), None, bytes([1]) + h[2].tobytes(), cachefmt.encode_pdqhash(h[3])]
    cv = [cachefmt.encode_coefficients(c[0]), None, cachefmt.encode_coefficients(c[2]), None]
    hashes, has_hash, coeffs, has_coeffs, q = cachefmt.load_cached(hv, cv, [80, None, 10, 49])
    assert has_hash.tolist() == [1, 0, 0, 1] and has_coeffs.tolist() == [1, 0, 1, 0]
    assert np.array_equal(hashes[0], h[0]) and np.array_equal(hashes[3], h[3]) and not hashes[2].any()
    assert np.array_equal(coeffs[0], c[0]) and q == [80, None, 10, 49]
    with pytest.raises(cachefmt.Corrupted):
        cachefmt.load_cached([hv[0]], [cachefmt.encode_coefficients(np.ones(5, np.float32))])
