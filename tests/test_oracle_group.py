"""Pins oracle/oracle_group.c and oracle_phash.c with the portable tests of
/root/reference/src/hamminghash.rs:273-412 and independent brute-force statements.  CPU only."""
import numpy as np
import pytest

from rupphash_b200.synth import planted_hashes, random_variants


def canon(groups):
    return sorted(sorted(g) for g in groups)


def test_high_similarity_support_u64(orc):
    """hamminghash.rs:287-307"""
    idx = orc.MIHIndex(np.array([0, 0xFFF], np.uint64))
    groups = idx.find_groups(12)
    assert len(groups) >= 1 and set(groups[0]) == {0, 1} and len(groups[0]) == 2


def test_high_similarity_support_pdq(orc):
    """hamminghash.rs:310-331"""
    h = np.zeros((2, 32), np.uint8)
    for i in range(30):
        h[1, i // 8] |= 1 << (i % 8)
    assert orc.hamming256(h[0], h[1]) == 30
    groups = orc.MIHIndex(h).find_groups(30)
    assert len(groups) >= 1 and {0, 1} <= set(groups[0])


def test_planted_cluster_u64(orc):
    """hamminghash.rs:336-412 at reduced n with a fixed seed (the reference uses an unseeded rng)"""
    n = 200_000
    rng = np.random.default_rng(1234)
    hashes = rng.integers(0, 2**63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    target = 0xABCD_1234_5678_90EF
    cluster = [target, target ^ 1, target ^ 2, target ^ 0x8000, target ^ 0x8001]
    inj = rng.choice(n, size=5, replace=False)
    for k, i in enumerate(inj):
        hashes[i] = np.uint64(cluster[k])
    groups = orc.MIHIndex(hashes).find_groups(5, threads=4)
    found = [g for g in groups if int(inj[0]) in g]
    assert found, "injected images not found"
    assert set(int(i) for i in inj) <= set(found[0])


def test_mih_index_csr(orc):
    """hamminghash.rs:89-138: bucket(k, chunk_k(h)) contains h's id, ids ascending, total = n*chunks"""
    rng = np.random.default_rng(2)
    h = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    idx = orc.MIHIndex(h)
    total = 0
    for k in range(16):
        seen = {}
        for i in range(500):
            v = int(h[i, 2 * k]) | (int(h[i, 2 * k + 1]) << 8)  # u16::from_le_bytes, hamminghash.rs:50-53
            seen.setdefault(v, []).append(i)
        for v, ids in seen.items():
            assert idx.bucket(k, v).tolist() == ids
            total += len(ids)
    assert total == 500 * 16


def brute_find_groups(h, max_dist, dist):
    """find_groups semantics (hamminghash.rs:245-270) over a complete adjacency"""
    n = len(h)
    adj = [[j for j in range(n) if j != i and dist(h[i], h[j]) <= max_dist] for i in range(n)]
    visited = [False] * n
    groups = []
    for i in range(n):
        if visited[i] or not adj[i]:
            continue
        g = [i]
        visited[i] = True
        for nb in adj[i]:
            if not visited[nb]:
                visited[nb] = True
                g.append(nb)
        if len(g) > 1:
            groups.append(g)
    return groups


@pytest.mark.parametrize("max_dist", [0, 5, 15, 31])
def test_find_groups_star_semantics_pdq(orc, max_dist):
    """complete for max_dist <= 31; groups compared as sets, seeds in order"""
    h, _ = planted_hashes(600, seed=21, threshold=31)
    got = orc.MIHIndex(h).find_groups(max_dist, threads=2)
    want = brute_find_groups(h, max_dist, orc.hamming256)
    assert [g[0] for g in got] == [g[0] for g in want]
    assert [sorted(g) for g in got] == [sorted(g) for g in want]


def brute_group(h, sim, has_hash=None, variants=None, n_variants=None, low_conf=None):
    """scanner.rs:1640-1817 edge rule, written directly from SURVEY 8a G4 with numpy popcounts"""
    n = len(h)
    bits = np.unpackbits(h, axis=1).astype(np.int16)
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    edges = []
    for i in range(n):
        if has_hash is not None and not has_hash[i]:
            continue
        if variants is None:
            vs = [bits[i]]
        else:
            cnt = 8 if n_variants is None else int(n_variants[i])
            vs = [np.unpackbits(variants[i, v]).astype(np.int16) for v in range(cnt)]
        for vb in vs:
            d = np.abs(bits[i + 1:] - vb[None, :]).sum(axis=1)
            for off in np.nonzero(d <= sim)[0]:
                j = i + 1 + int(off)
                if has_hash is not None and not has_hash[j]:
                    continue
                lim = 0 if (low_conf is not None and (low_conf[i] or low_conf[j])) else sim
                if d[off] <= lim:
                    edges.append((i, j))
                    a, b = find(i), find(j)
                    if a != b:
                        parent[max(a, b)] = min(a, b)
    labels = np.array([find(i) for i in range(n)], np.uint32)
    return labels, edges


@pytest.mark.parametrize("sim", [0, 15, 31, 40, 63])
@pytest.mark.parametrize("mode", ["plain", "lowconf", "variants", "holes"])
def test_group_generic_matches_brute_force(orc, sim, mode):
    n = 700
    h, lc = planted_hashes(n, seed=100 + sim, threshold=min(sim, 40))
    kw = {}
    if mode in ("lowconf", "variants", "holes"):
        kw["low_conf"] = lc
    if mode in ("variants", "holes"):
        kw["variants"] = random_variants(h, seed=sim)
        nv = np.full(n, 8, np.uint8)
        nv[::7] = 1
        nv[3::11] = 4
        kw["n_variants"] = nv
    if mode == "holes":
        hh = np.ones(n, np.uint8)
        hh[::5] = 0
        kw["has_hash"] = hh
    want_labels, want_edges = brute_group(h, sim, **kw)
    for use_mih in (True, False):
        labels, count, edges = orc.group_generic(h, sim, threads=3, use_mih=use_mih, edges_cap=200000, **kw)
        assert count == len(want_edges)
        assert sorted(map(tuple, edges.tolist())) == sorted(want_edges)
        assert np.array_equal(labels, want_labels)


def test_group_generic_rejects_similarity_above_63(orc):
    """assert at scanner.rs:1650-1655"""
    h, _ = planted_hashes(100)
    with pytest.raises(ValueError):
        orc.group_generic(h, 64)


def test_group_generic_empty_and_single(orc):
    labels, count, _ = orc.group_generic(np.zeros((0, 32), np.uint8), 31)
    assert labels.size == 0 and count == 0
    labels, count, _ = orc.group_generic(np.zeros((1, 32), np.uint8), 31)
    assert labels.tolist() == [0] and count == 0
    h = np.zeros((3, 32), np.uint8)
    labels, count, _ = orc.group_generic(h, 0, has_hash=np.array([0, 0, 0], np.uint8))
    assert labels.tolist() == [0, 1, 2] and count == 0


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("with_variants", [False, True])
def test_tile_ranks_merge_to_the_same_labels(orc, world, with_variants):
    n = 900
    h, lc = planted_hashes(n, seed=77)
    kw = {"low_conf": lc}
    if with_variants:
        kw["variants"] = random_variants(h, seed=3)
        nv = np.full(n, 8, np.uint8)
        nv[::3] = 2
        kw["n_variants"] = nv
        hh = np.ones(n, np.uint8)
        hh[10:40] = 0
        kw["has_hash"] = hh
    want, want_count, _ = orc.group_generic(h, 31, **kw)
    parents, total = [], 0
    for r in range(world):
        p, c = orc.group_tiles_rank(h, 31, 128, r, world, **kw)
        parents.append(p)
        total += c
    assert total == want_count
    assert np.array_equal(orc.merge_parents(np.stack(parents)), want)


def test_labels_to_groups(orc):
    labels = np.array([0, 1, 0, 3, 1, 5], np.uint32)
    assert orc.labels_to_groups(labels) == [[0, 2], [1, 4]]


# ------------------------------------------------------------------ pHash bit ops (phash.rs:137-255)

def py_rot(hash_, kind):
    res = 0
    for y in range(8):
        for x in range(8):
            src = 8 * y + x
            bit = (hash_ >> (63 - src)) & 1
            if kind in (90, 270):
                dx, dy = y, x
                flip = (dx % 2 != 0) if kind == 90 else (dy % 2 != 0)
            elif kind == 180:
                dx, dy = x, y
                flip = (x + y) % 2 != 0
            else:  # flip_h
                dx, dy = x, y
                flip = x % 2 != 0
            res |= (bit ^ int(flip)) << (63 - (8 * dy + dx))
    return res


def test_phash_bit_ops(orc):
    rng = np.random.default_rng(9)
    for h in [0, 2**64 - 1, 0xDEB1E20C136F983C] + [int(x) for x in rng.integers(0, 2**63, 200, dtype=np.uint64)]:
        assert orc.phash_rot90(h) == py_rot(h, 90)
        assert orc.phash_rot180(h) == py_rot(h, 180)
        assert orc.phash_rot270(h) == py_rot(h, 270)
        assert orc.phash_flip_h(h) == py_rot(h, "f")
        f = py_rot(h, "f")
        assert orc.phash_dihedral(h) == [h, py_rot(h, 90), py_rot(h, 180), py_rot(h, 270),
                                         f, py_rot(f, 90), py_rot(f, 180), py_rot(f, 270)]
        assert orc.phash_rot_invariant(h) == min(h, py_rot(h, 90), py_rot(h, 180), py_rot(h, 270))


def test_phash_from_luma32_matches_float64_dct(orc):
    """bits agree with a float64 DCT except where a coefficient is within 1e-3 of the median"""
    from scipy.fft import dctn

    rng = np.random.default_rng(10)
    for _ in range(20):
        luma = rng.integers(0, 256, (32, 32), dtype=np.uint8)
        got = orc.phash_from_luma32(luma)
        coef = dctn(luma.astype(np.float64), type=2, norm=None)[:8, :8] / 4.0
        flat = coef.reshape(64)
        median = np.sort(flat[1:])[31]
        for i in range(64):
            if abs(flat[i] - median) > 1e-3 * max(1.0, abs(median)):
                assert ((got >> (63 - i)) & 1) == int(flat[i] > median)
