"""SURVEY 8f N3 on the device: a library grouped from the reference's cache entries alone (hash_db /
coeff_db plaintext values + the stored quality short) equals the grouping of the freshly hashed
pixels, and both equal the CPU oracle."""
import numpy as np
import pytest

from rupphash_b200.synth import synth_images

pytestmark = pytest.mark.gpu


def test_regroup_from_cache_equals_fresh_grouping(orc):
    from rupphash_b200 import _lib, cachefmt, pdqhash, scanner
    ctx = _lib.Context(0)
    base = synth_images(40, 384, 512, seed=21)
    # near-duplicates (noise), rotated / flipped copies (found only through the dihedral variants), a flat image
    rng = np.random.default_rng(5)
    noisy = np.clip(base[:10].astype(np.int16) + rng.integers(-3, 4, size=base[:10].shape), 0, 255).astype(np.uint8)
    flipped = base[10:20, :, ::-1].copy()
    rot180 = base[20:25, ::-1, ::-1].copy()
    flat = np.full((1, 384, 512, 3), 77, np.uint8)
    imgs = np.concatenate([base, noisy, flipped, rot180, flat])
    out = pdqhash.hash_batch(imgs, want_coeffs=True, ctx=ctx)
    rows = cachefmt.encode_batch(out)
    hv, cv, q = [r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows]
    # a few files lost their coefficients, one has an entry from the old pipeline, one has no quality
    for k in (3, 41, 55):
        cv[k] = None
    cv[7] = bytes([1]) + cv[7][1:]
    q[12] = None
    groups, cnt = scanner.regroup_from_cache(hv, cv, q, 31, ctx=ctx)

    # the same through the arrays, and through the oracle
    hashes, has_hash, coeffs, has_coeffs, q2 = cachefmt.load_cached(hv, cv, q)
    assert np.array_equal(hashes, out["hash"]) and has_hash.all()
    assert has_coeffs.sum() == len(imgs) - 4
    variants = np.zeros((len(imgs), 8, 32), np.uint8)
    variants[:, 0] = hashes
    nv = np.ones(len(imgs), np.uint8)
    idx = np.flatnonzero(has_coeffs)
    variants[idx] = np.stack([orc.dihedral(coeffs[i]) for i in idx])
    nv[idx] = 8
    low = np.array([scanner.is_low_confidence(v) for v in q2], np.uint8)
    ref_labels, ref_cnt, _ = orc.group_generic(hashes, 31, variants=variants, n_variants=nv, low_conf=low, use_mih=False)
    assert cnt == ref_cnt
    assert groups == orc.labels_to_groups(ref_labels)
    # the planted relations are found: noisy copies join their originals, flipped / rotated ones too
    # (unless their coefficients are gone and the original's are too)
    member = {i: g for g in groups for i in g}
    assert all(member.get(k) is not None and 40 + k in member[k] for k in range(10))
    assert sum(50 + k in member.get(10 + k, []) for k in range(10)) >= 8
    assert low[-1] == 1   # the flat image has quality 0: low confidence, grouped by exact match only
    ctx.close()
