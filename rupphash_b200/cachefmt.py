"""Cache wire formats of the reference's LMDB stores, so that device results drop
into a cache written by `phdupes` unchanged and a cached library can be re-grouped on the GPU without
re-hashing.  These are the PLAINTEXT payloads; the XChaCha20-Poly1305 envelope around them
(db.rs:640-673) and LMDB itself stay with the reference.

  hash_db   value = [PDQ_ALGO_VERSION] + 32-byte hash                      db.rs:1200-1210, read :678-702
  coeff_db  value = [PDQ_ALGO_VERSION] + postcard(CachedCoefficients)      db.rs:1221-1231, read :734-761
            postcard of `struct { coefficients: Vec<f32> }` = LEB128 varint length + f32 little-endian
            each (postcard 1.x: seq = varint(usize) len + elements; f32 = 4 LE bytes)
  quality   ImageFeatures tag TAG_DERIVED_PDQ_QUALITY (0xF007) = ExifValue::Short(q100)
            scanner.rs:1416-1418, :1465-1473; exif_types.rs:74; image_features.rs:107-112
"""
from __future__ import annotations

import numpy as np

PDQ_ALGO_VERSION = 2               # db.rs:47 ("2 = reference-compatible PDQ")
TAG_DERIVED_PDQ_QUALITY = 0xF007   # exif_types.rs:74
N_COEFFS = 256


class Corrupted(ValueError):
    """lmdb::Error::Corrupted of the reference's readers (db.rs:696, :747, :753)."""


def _varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf: bytes, pos: int):
    shift = val = 0
    while True:
        if pos >= len(buf) or shift > 63:
            raise Corrupted("truncated varint")
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def encode_pdqhash(hash32) -> bytes:
    """db.rs:1203-1206"""
    h = np.ascontiguousarray(hash32, dtype=np.uint8).reshape(-1)
    if h.size != 32:
        raise ValueError("PDQ hash is 32 bytes")
    return bytes([PDQ_ALGO_VERSION]) + h.tobytes()


def decode_pdqhash(value: bytes):
    """db.rs:686-694: another version or length is a MISS (None), not an error."""
    if len(value) == 33 and value[0] == PDQ_ALGO_VERSION:
        return np.frombuffer(value, dtype=np.uint8, count=32, offset=1).copy()
    return None


def encode_coefficients(coeffs) -> bytes:
    """db.rs:1224-1228 + CachedCoefficients::to_bytes (db.rs:224-226)"""
    c = np.ascontiguousarray(coeffs, dtype="<f4").reshape(-1)
    return bytes([PDQ_ALGO_VERSION]) + _varint(c.size) + c.tobytes()


def decode_coefficients(value: bytes):
    """db.rs:742-754: older version -> None (absent); right version but undecodable -> Corrupted."""
    if len(value) == 0 or value[0] != PDQ_ALGO_VERSION:
        return None
    n, pos = _read_varint(value, 1)
    if len(value) - pos != 4 * n:
        raise Corrupted("coefficient payload length does not match its postcard length prefix")
    return np.frombuffer(value, dtype="<f4", count=n, offset=pos).astype(np.float32)


def quality_tag(quality: float):
    """(tag id, u16 value) stored as ExifValue::Short (scanner.rs:1416-1418, :1470-1473)."""
    from .scanner import quality_100
    return TAG_DERIVED_PDQ_QUALITY, quality_100(quality)


def encode_batch(out: dict):
    """hash_batch(..., want_coeffs=True) output -> per-image (hash_db value, coeff_db value or None,
    q100 or None); invalid images (pdqhash.rs:167-169) produce no entries, as in the scanner."""
    from .scanner import quality_100
    hashes, quality, valid = np.asarray(out["hash"]), np.asarray(out["quality"]), np.asarray(out["valid"])
    coeffs = None if out.get("coeffs") is None else np.asarray(out["coeffs"])
    rows = []
    for i in range(len(hashes)):
        if not valid[i]:
            rows.append((None, None, None))
            continue
        rows.append((encode_pdqhash(hashes[i]), None if coeffs is None else encode_coefficients(coeffs[i]),
                     quality_100(float(quality[i]))))
    return rows


def load_cached(hash_values, coeff_values=None, q100=None):
    """Cached values (bytes or None per file) -> the arrays group_with_pdqhash takes: hashes (n, 32),
    has_hash (n,), coeffs (n, 256) with has_coeffs (n,), q100 list.  Entries of another algorithm
    version count as absent, exactly as the reference's readers treat them."""
    n = len(hash_values)
    hashes = np.zeros((n, 32), np.uint8)
    has_hash = np.zeros(n, np.uint8)
    coeffs = np.zeros((n, N_COEFFS), np.float32)
    has_coeffs = np.zeros(n, np.uint8)
    for i, v in enumerate(hash_values):
        h = None if v is None else decode_pdqhash(v)
        if h is not None:
            hashes[i] = h
            has_hash[i] = 1
    if coeff_values is not None:
        for i, v in enumerate(coeff_values):
            c = None if v is None else decode_coefficients(v)
            if c is not None:
                if c.size != N_COEFFS:
                    raise Corrupted("PDQ coefficient vectors have 256 entries")
                coeffs[i] = c
                has_coeffs[i] = 1
    return hashes, has_hash, coeffs, has_coeffs, (list(q100) if q100 is not None else [None] * n)
