"""Group post-processing: the host mirror in rupphash_b200/scanner.py (key-based sorts, one
batched max_dist reduce) against the oracle's statement-by-statement restatement of scanner.rs:1986-2022,
:2040-2110, :2183-2262 and :1561-1576 (oracle/postprocess.py, comparator-based, per-group loops) on a planted
library with paths.  CPU part: the device calls are replaced by the oracle's distance code; the gpu-marked
test runs the same library through the device."""
import hashlib

import numpy as np
import pytest

from rupphash_b200 import scanner


def planted_library(orc, seed=5, n_groups=60):
    """Files with paths, sizes, dates, content hashes, pixel hashes, PDQ hashes and (for most) coefficients."""
    rng = np.random.default_rng(seed)
    files, coeffs, raw_groups = [], {}, []
    exts = ["jpg", "JPG", "png", "cr2", "NEF", "jpeg", "dng", "tif"]
    for g in range(n_groups):
        k = int(rng.integers(2, 7))
        base_c = (rng.standard_normal(256) * 30).astype(np.float32)
        idxs = []
        dup_content = hashlib.sha256(f"dup{g}".encode()).digest()
        for m in range(k):
            i = len(files)
            c = base_c + (rng.standard_normal(256) * 2).astype(np.float32)
            has_feat = rng.random() > 0.25
            has_hash = has_feat or rng.random() > 0.3
            identical = g % 3 == 0 and m < 2 or g % 7 == 0           # some / all members bit-identical
            content = dup_content if identical else hashlib.sha256(f"{g}-{m}".encode()).digest()
            stem = ["IMG_%d" % (g * 3 + (m // 2)), "img%d" % int(rng.integers(1, 120)), "a %d b" % m, "x0%d" % m][int(rng.integers(0, 4))]
            path = f"/lib/dir{g % 5}/g{g}/{stem}.{exts[m % len(exts)] if m % 2 else exts[int(rng.integers(0, len(exts)))]}"
            while any(f.path == path for f in files):           # paths are unique in a scan (canonicalised file names)
                path = path.replace("/g%d/" % g, "/g%d_/" % g)
            f = scanner.FileMeta(path=path, size=int(rng.integers(1000, 10**7)), modified=int(rng.integers(10**9, 2 * 10**9)),
                                 content_hash=content,
                                 pixel_hash=hashlib.sha256(content).digest() if rng.random() > 0.5 else None,
                                 pdqhash=bytes(orc.to_hash(c)) if has_hash else None,
                                 exif_timestamp=int(rng.integers(10**9, 2 * 10**9)) if rng.random() > 0.5 else None, index=i)
            files.append(f)
            if has_feat:
                coeffs[i] = c
            idxs.append(i)
        rng.shuffle(idxs)
        raw_groups.append([int(x) for x in idxs])
    return files, coeffs, raw_groups


def oracle_reduce(orc):
    def max_dist(piv, nv, mem_h, mem_g, ng):
        out = np.zeros(ng, np.uint32)
        for h, g in zip(mem_h, mem_g):
            d = min(orc.hamming256(piv[g, v], h) for v in range(int(nv[g])))
            out[g] = max(out[g], d)
        return out
    return max_dist


def as_rows(groups, infos):
    return [([f.path for f in g], i["max_dist"], i["status"]) for g, i in zip(groups, infos)]


@pytest.mark.parametrize("sort_order", ["name-natural", "name", "name-desc", "date", "date-desc", "size", "size-desc",
                                        "exif-date", "exif-date-desc", "name-natural-desc", "unknown-falls-back"])
def test_mirror_equals_restatement(orc, sort_order):
    from oracle import postprocess
    files, coeffs, raw = planted_library(orc)
    got_g, got_i = scanner.process_raw_groups(raw, files, sort_order, coeffs, max_dist_fn=oracle_reduce(orc),
                                              dihedral_fn=lambda c: np.stack([orc.dihedral(x) for x in c]))
    got_g, got_i = scanner.sort_groups(got_g, got_i)
    features = {files[i].path: c for i, c in coeffs.items()}
    want_g, want_i = postprocess.process_and_sort(raw, files, features, sort_order, orc)
    assert as_rows(got_g, got_i) == as_rows(want_g, want_i)
    statuses = {i["status"] for i in got_i}
    assert statuses == {"None", "SomeIdentical", "AllIdentical"}
    flags = [i["status"] != "None" for i in got_i]
    assert flags == sorted(flags, reverse=True)                     # groups with identical files first (scanner.rs:1563-1567)


def test_natural_order_matches_strnatcmp_cases():
    from oracle import postprocess
    import functools
    names = ["img12.jpg", "img10.jpg", "img2.jpg", "img02.jpg", "img1.jpg", "IMG1.jpg", "a 5.png", "a5.png", "x007", "x7",
             "x0070", "pic 3", "pic3a", "1.5", "1.10", "1.05", ""]
    want = sorted(names, key=functools.cmp_to_key(postprocess.natord_compare))
    assert sorted(names, key=scanner.natural_key) == want
    assert want.index("img2.jpg") < want.index("img10.jpg") < want.index("img12.jpg")
    assert want.index("x007") < want.index("x7")                    # a leading zero compares like a fraction


def test_raw_extension_and_stem_rules():
    assert scanner.is_raw_ext("/a/b/IMG_1.CR2") and scanner.is_raw_ext("x.nef") and not scanner.is_raw_ext("x.jpg")
    assert not scanner.is_raw_ext("/a/.cr2") and not scanner.is_raw_ext("/a/noext")
    fs = [scanner.FileMeta("/d/IMG_2.cr2"), scanner.FileMeta("/d/IMG_2.jpg"), scanner.FileMeta("/d/IMG_1.dng"),
          scanner.FileMeta("/e/IMG_1.png")]
    scanner.sort_by_stem_then_ext(fs)
    assert [f.path for f in fs] == ["/e/IMG_1.png", "/d/IMG_1.dng", "/d/IMG_2.jpg", "/d/IMG_2.cr2"]


def test_group_without_any_hash_has_zero_max_dist(orc):
    files = [scanner.FileMeta("/a/x.jpg", content_hash=b"1"), scanner.FileMeta("/a/y.jpg", content_hash=b"2")]
    g, i = scanner.process_raw_groups([[0, 1]], files, "name", {}, max_dist_fn=oracle_reduce(orc))
    assert i == [{"max_dist": 0, "status": "None"}] and [f.path for f in g[0]] == ["/a/x.jpg", "/a/y.jpg"]


@pytest.mark.gpu
def test_device_pipeline_prints_what_the_reference_prints(orc):
    """scan_groups: device edge phase + union-find, stem merge, ordering, device max_dist, final order -- against
    the oracle's grouping + post-processing on the same library."""
    from oracle import postprocess
    from rupphash_b200 import _lib
    ctx = _lib.Context(0)
    try:
        files, coeffs, _ = planted_library(orc, seed=9, n_groups=80)
        q100 = [None if k % 9 == 0 else (30 if k % 11 == 0 else 90) for k in range(len(files))]
        groups, infos, comparisons = scanner.scan_groups(files, 40, "name-natural", coeffs, q100, ctx=ctx)
        n = len(files)
        hashes = np.zeros((n, 32), np.uint8)
        has_hash = np.zeros(n, np.uint8)
        variants = np.zeros((n, 8, 32), np.uint8)
        nv = np.ones(n, np.uint8)
        for i, f in enumerate(files):
            if f.pdqhash is not None:
                hashes[i] = np.frombuffer(f.pdqhash, np.uint8)
                has_hash[i] = 1
            variants[i, 0] = hashes[i]
            if i in coeffs and has_hash[i]:
                variants[i] = orc.dihedral(coeffs[i])
                nv[i] = 8
        low = np.array([1 if (q is not None and q < 50) else 0 for q in q100], np.uint8)
        labels, cnt, _ = orc.group_generic(hashes, 40, has_hash=has_hash, variants=variants, n_variants=nv, low_conf=low)
        assert comparisons == cnt
        raw = scanner.merge_groups_by_stem(orc.labels_to_groups(labels), [f.path for f in files])
        features = {files[i].path: c for i, c in coeffs.items()}
        want_g, want_i = postprocess.process_and_sort(raw, files, features, "name-natural", orc)
        assert as_rows(groups, infos) == as_rows(want_g, want_i)
        assert len(groups) > 10
    finally:
        ctx.close()
