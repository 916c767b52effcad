"""CPU restatement of the reference's group post-processing, TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows the Rust statement by statement -- comparator-based sorts, per-group loops, distances through the
oracle's own hamming / dihedral code -- so that the product's key-based, batched mirror
(rupphash_b200/scanner.py process_raw_groups / sort_groups) is checked against an independent statement:

    analyze_group_with_features   scanner.rs:2183-2254
    sort_files                    scanner.rs:2040-2110
    sort_by_stem_then_ext         scanner.rs:2256-2262
    process_raw_groups            scanner.rs:1986-2022
    final group order             scanner.rs:1561-1576
    natord::compare               crate `natord` 1.0.9 (Cargo.lock), not in the reference tree: restated
                                  from its published algorithm (strnatcmp); parity with the crate unpinned
"""
from __future__ import annotations

import functools
import os

import numpy as np

RAW_EXTS = {"nef", "dng", "cr2", "cr3", "arw", "orf", "rw2", "raf", "kdc", "dcr", "pef", "x3f", "srf", "3fr"}  # scanner.rs:43-46


def _digits_right(a, ai, b, bi):
    """natord compare_right: the longest run of digits wins; equal length: the first difference."""
    bias = 0
    while True:
        ca = a[ai] if ai < len(a) else None
        cb = b[bi] if bi < len(b) else None
        da, db = ca is not None and ca.isdigit(), cb is not None and cb.isdigit()
        if not da and not db:
            return bias, ai, bi
        if not da:
            return -1, ai, bi
        if not db:
            return 1, ai, bi
        if ca < cb and bias == 0:
            bias = -1
        elif ca > cb and bias == 0:
            bias = 1
        ai, bi = ai + 1, bi + 1


def _digits_left(a, ai, b, bi):
    """natord compare_left: digit by digit, left aligned."""
    while True:
        ca = a[ai] if ai < len(a) else None
        cb = b[bi] if bi < len(b) else None
        da, db = ca is not None and ca.isdigit(), cb is not None and cb.isdigit()
        if not da and not db:
            return 0, ai, bi
        if not da:
            return -1, ai, bi
        if not db:
            return 1, ai, bi
        if ca < cb:
            return -1, ai, bi
        if ca > cb:
            return 1, ai, bi
        ai, bi = ai + 1, bi + 1


def natord_compare(a: str, b: str) -> int:
    ai = bi = 0
    while True:
        while ai < len(a) and a[ai].isspace():
            ai += 1
        while bi < len(b) and b[bi].isspace():
            bi += 1
        ca = a[ai] if ai < len(a) else None
        cb = b[bi] if bi < len(b) else None
        if ca is not None and cb is not None and ca.isdigit() and cb.isdigit():
            fractional = ca == "0" or cb == "0"
            r, ai, bi = (_digits_left if fractional else _digits_right)(a, ai, b, bi)
            if r != 0:
                return r
            continue
        if ca is None and cb is None:
            return 0
        if ca is None:
            return -1
        if cb is None:
            return 1
        if ca != cb:
            return -1 if ca < cb else 1
        ai, bi = ai + 1, bi + 1


def _cmp(x, y):
    return (x > y) - (x < y)


def _name(f):
    return os.path.basename(f.path)


def sort_files(files, sort_order):
    """scanner.rs:2040-2110 (stable sorts)"""
    nat = functools.cmp_to_key(lambda x, y: natord_compare(_name(x), _name(y)))
    if sort_order == "name":
        files.sort(key=functools.cmp_to_key(lambda x, y: _cmp(_name(x), _name(y))))
    elif sort_order == "name-desc":
        files.sort(key=functools.cmp_to_key(lambda x, y: _cmp(_name(x), _name(y))))
        files.reverse()
    elif sort_order == "name-natural":
        files.sort(key=nat)
    elif sort_order == "name-natural-desc":
        files.sort(key=nat)
        files.reverse()
    elif sort_order == "date":
        files.sort(key=functools.cmp_to_key(lambda x, y: _cmp(x.modified, y.modified)))
    elif sort_order == "date-desc":
        files.sort(key=functools.cmp_to_key(lambda x, y: _cmp(y.modified, x.modified)))
    elif sort_order == "size":
        files.sort(key=functools.cmp_to_key(lambda x, y: _cmp(x.size, y.size)))
    elif sort_order == "size-desc":
        files.sort(key=functools.cmp_to_key(lambda x, y: _cmp(y.size, x.size)))
    elif sort_order in ("exif-date", "exif-date-desc"):
        desc = sort_order.endswith("desc")

        def c(x, y):
            tx, ty = x.exif_timestamp, y.exif_timestamp
            if tx is not None and ty is not None:
                return _cmp(ty, tx) if desc else _cmp(tx, ty)
            if tx is not None:
                return -1
            if ty is not None:
                return 1
            return _cmp(y.modified, x.modified) if desc else _cmp(x.modified, y.modified)
        files.sort(key=functools.cmp_to_key(c))
    elif sort_order in ("random", "location"):
        return
    else:
        files.sort(key=nat)


def is_raw_ext(path):
    base = os.path.basename(path)
    if "." not in base.lstrip(".") and not (base.count(".") and not base.startswith(".")):
        return False
    ext = base.rsplit(".", 1)[1] if "." in base[1:] else ""
    return ext.lower() in RAW_EXTS


def sort_by_stem_then_ext(files):
    """scanner.rs:2256-2262"""
    def stem(f):
        base = os.path.basename(f.path)
        if "." in base[1:]:
            return base[: base.rindex(".")]
        return base

    def c(x, y):
        r = _cmp(stem(x), stem(y))
        return r if r else _cmp(is_raw_ext(x.path), is_raw_ext(y.path))
    files.sort(key=functools.cmp_to_key(c))


def analyze_group_with_features(files, features, sort_order, orc):
    """scanner.rs:2183-2254 -> (files in display order, max_dist, status).  features: {path: 256 f32}."""
    if not files:
        return [], 0, "None"
    counts = {}
    for f in files:
        counts[f.content_hash] = counts.get(f.content_hash, 0) + 1
    duplicates = [f for f in files if counts.get(f.content_hash, 0) > 1]
    unique = [f for f in files if not counts.get(f.content_hash, 0) > 1]

    def dup_cmp(x, y):
        for kx, ky in (((x.pixel_hash is not None, bytes(x.pixel_hash or b"")), (y.pixel_hash is not None, bytes(y.pixel_hash or b""))),
                       (bytes(x.content_hash), bytes(y.content_hash)), (_name(x), _name(y))):
            r = _cmp(kx, ky)
            if r:
                return r
        return 0
    duplicates.sort(key=functools.cmp_to_key(dup_cmp))
    sort_files(unique, sort_order)
    files = duplicates + unique
    sort_by_stem_then_ext(files)
    pivot_feats = next((features[f.path] for f in files if f.path in features), None)
    hashes = [np.frombuffer(bytes(f.pdqhash), np.uint8) for f in files if f.pdqhash is not None]
    if pivot_feats is not None:
        variants = orc.dihedral(np.asarray(pivot_feats, np.float32))
        max_d = max((min(orc.hamming256(v, h) for v in variants) for h in hashes), default=0)
    elif hashes:
        pivot = hashes[0]
        max_d = max(orc.hamming256(pivot, h) for h in hashes)
    else:
        max_d = 0
    if len(counts) == 1:
        status = "AllIdentical"
    elif not all(c == 1 for c in counts.values()):
        status = "SomeIdentical"
    else:
        status = "None"
    return files, max_d, status


def process_and_sort(raw_groups, files, features, sort_order, orc):
    """process_raw_groups (scanner.rs:1986-2022) + the final order (scanner.rs:1561-1576)."""
    out = []
    for idxs in raw_groups:
        arranged, max_d, status = analyze_group_with_features([files[i] for i in idxs], features, sort_order.lower(), orc)
        out.append((arranged, {"max_dist": max_d, "status": status}))

    def c(a, b):
        (g1, i1), (g2, i2) = a, b
        h1, h2 = i1["status"] != "None", i2["status"] != "None"
        if h1 != h2:
            return _cmp(h2, h1)
        if i1["max_dist"] != i2["max_dist"]:
            return _cmp(i1["max_dist"], i2["max_dist"])
        s1 = g1[0].size if g1 else 0
        s2 = g2[0].size if g2 else 0
        return _cmp(s2, s1)
    out.sort(key=functools.cmp_to_key(c))
    return [g for g, _ in out], [i for _, i in out]
