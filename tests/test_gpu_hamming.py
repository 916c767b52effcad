"""GPU parity tests of hot path #2 (all-pairs Hamming + union-find) through the C ABI, against
the CPU oracle on the same seeded inputs.  Bar: bit-exact labels, edge counts and distances."""
import numpy as np
import pytest

from rupphash_b200.synth import planted_hashes, random_variants

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from rupphash_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def test_distances_bit_exact(ctx, orc):
    from rupphash_b200 import hamminghash
    rng = np.random.default_rng(0)
    n = 100_000
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    b = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[0] = 0; b[0] = 0            # distance 0
    a[1] = 0; b[1] = 255          # distance 256
    a[2] = 0; b[2] = 0; b[2, 31] = 1  # single bit
    b[3] = a[3]
    got = hamminghash.hamming_distances(a, b, ctx)
    x = np.bitwise_xor(a, b)
    want = np.unpackbits(x, axis=1).sum(axis=1).astype(np.uint32)
    assert np.array_equal(got, want)
    assert got[0] == 0 and got[1] == 256 and got[2] == 1 and got[3] == 0
    for i in range(0, 50):
        assert orc.hamming256(a[i], b[i]) == got[i]
    u = rng.integers(0, 2**63, size=1000, dtype=np.uint64)
    v = rng.integers(0, 2**63, size=1000, dtype=np.uint64)
    got64 = hamminghash.hamming_distances(u, v, ctx)
    want64 = np.array([bin(int(p) ^ int(q)).count("1") for p, q in zip(u, v)], np.uint32)
    assert np.array_equal(got64, want64)
    assert hamminghash.hamming_distance(a[5], b[5], ctx) == want[5]
    assert hamminghash.hamming_distance(int(u[0]), int(v[0]), ctx) == want64[0]


@pytest.mark.parametrize("similarity", [0, 15, 31, 40, 63])
@pytest.mark.parametrize("use_variants", [False, True])
@pytest.mark.parametrize("use_low_conf", [False, True])
def test_group_small_vs_bruteforce(ctx, orc, similarity, use_variants, use_low_conf):
    from rupphash_b200 import scanner
    n = 3000
    hashes, low_conf = planted_hashes(n, seed=100 + similarity, threshold=max(similarity, 1))
    variants = random_variants(hashes, seed=3) if use_variants else None
    lc = low_conf if use_low_conf else None
    labels, cnt = scanner.group_labels(hashes, similarity, variants=variants, low_conf=lc, ctx=ctx)
    ref_labels, ref_cnt, ref_edges = orc.group_generic(hashes, similarity, variants=variants, low_conf=lc,
                                                       use_mih=False, edges_cap=2_000_000)
    assert cnt == ref_cnt
    assert np.array_equal(labels, ref_labels)
    got_edges, cnt2 = scanner.edges(hashes, similarity, variants=variants, low_conf=lc, cap=2_000_000, ctx=ctx)
    assert cnt2 == ref_cnt and len(got_edges) == ref_cnt
    key = lambda e: np.sort(e[:, 0].astype(np.uint64) << np.uint64(32) | e[:, 1].astype(np.uint64))
    assert np.array_equal(key(got_edges), key(ref_edges)), "edge multiset differs"
    assert scanner.labels_to_groups(labels) == orc.labels_to_groups(ref_labels)


def test_group_has_hash_and_n_variants(ctx, orc):
    from rupphash_b200 import scanner
    n = 5000
    rng = np.random.default_rng(9)
    hashes, low_conf = planted_hashes(n, seed=77)
    variants = random_variants(hashes, seed=5)
    has_hash = (rng.random(n) > 0.2).astype(np.uint8)
    n_variants = rng.choice(np.array([0, 1, 3, 8], np.uint8), size=n)
    labels, cnt = scanner.group_labels(hashes, 31, has_hash=has_hash, variants=variants, n_variants=n_variants,
                                       low_conf=low_conf, ctx=ctx)
    ref_labels, ref_cnt, _ = orc.group_generic(hashes, 31, has_hash=has_hash, variants=variants,
                                               n_variants=n_variants, low_conf=low_conf, use_mih=True)
    assert cnt == ref_cnt and np.array_equal(labels, ref_labels)
    # files without a hash are their own label
    assert np.array_equal(labels[has_hash == 0], np.flatnonzero(has_hash == 0))


@pytest.mark.parametrize("n", [1, 2, 63, 1023, 1024, 1025, 2049])
def test_group_ragged_sizes(ctx, orc, n):
    from rupphash_b200 import scanner
    hashes, low_conf = planted_hashes(n, seed=n)
    if n >= 2:
        hashes[n - 1] = hashes[0]  # an edge across the whole index range
    labels, cnt = scanner.group_labels(hashes, 31, low_conf=low_conf, ctx=ctx)
    ref_labels, ref_cnt, _ = orc.group_generic(hashes, 31, low_conf=low_conf, use_mih=False)
    assert cnt == ref_cnt and np.array_equal(labels, ref_labels)


def test_group_empty_and_errors(ctx):
    from rupphash_b200 import scanner
    labels, cnt = scanner.group_labels(np.zeros((0, 32), np.uint8), 31, ctx=ctx)
    assert len(labels) == 0 and cnt == 0
    with pytest.raises(ValueError):  # scanner.rs:1650-1655 assert
        scanner.group_labels(np.zeros((4, 32), np.uint8), 64, ctx=ctx)


def test_group_all_identical_hashes(ctx, orc):
    """Output explosion case: k identical hashes = k(k-1)/2 edges, one group."""
    from rupphash_b200 import scanner
    n = 1500
    hashes = np.tile(np.arange(32, dtype=np.uint8), (n, 1))
    labels, cnt = scanner.group_labels(hashes, 0, ctx=ctx)
    assert cnt == n * (n - 1) // 2
    assert np.all(labels == 0)
    lc = np.ones(n, np.uint8)
    labels, cnt = scanner.group_labels(hashes, 31, low_conf=lc, ctx=ctx)
    assert cnt == n * (n - 1) // 2 and np.all(labels == 0)


def test_group_50k_vs_mih_oracle(ctx, orc):
    from rupphash_b200 import scanner
    n = 50_000
    hashes, low_conf = planted_hashes(n, seed=0xB200)
    for variants in (None, random_variants(hashes, seed=1)):
        labels, cnt = scanner.group_labels(hashes, 31, variants=variants, low_conf=low_conf, ctx=ctx)
        ref_labels, ref_cnt, _ = orc.group_generic(hashes, 31, variants=variants, low_conf=low_conf, threads=8,
                                                   use_mih=True)
        assert cnt == ref_cnt and np.array_equal(labels, ref_labels)


def test_group_device_resident_inputs(ctx, orc):
    import torch
    from rupphash_b200 import scanner
    n = 20_000
    hashes, low_conf = planted_hashes(n, seed=21)
    dh = torch.from_numpy(hashes).cuda()
    dl = torch.from_numpy(low_conf).cuda()
    labels, cnt = scanner.group_labels(dh, 31, low_conf=dl, ctx=ctx)
    assert labels.is_cuda
    ref_labels, ref_cnt, _ = orc.group_generic(hashes, 31, low_conf=low_conf, threads=4)
    assert cnt == ref_cnt and np.array_equal(labels.cpu().numpy().view(np.uint32), ref_labels)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_shards_merge_to_single_gpu_result(ctx, orc, world):
    """Every rank's tile share run one after the other on one GPU, merged with rh_uf_merge:
    identical labels and edge count to the unsharded search (SURVEY 8e determinism)."""
    from rupphash_b200 import scanner
    n = 12_000
    hashes, low_conf = planted_hashes(n, seed=5)
    variants = random_variants(hashes, seed=8)
    labels, cnt = scanner.group_labels(hashes, 31, variants=variants, low_conf=low_conf, ctx=ctx)
    forests, total = [], 0
    for rank in range(world):
        p, c = scanner.group_shard(hashes, 31, rank, world, variants=variants, low_conf=low_conf, ctx=ctx)
        forests.append(np.asarray(p).copy())
        total += c
    merged = scanner.merge_forests(np.stack(forests), ctx)
    assert total == cnt
    assert np.array_equal(merged, labels)
    ref_labels, ref_cnt, _ = orc.group_generic(hashes, 31, variants=variants, low_conf=low_conf, threads=4)
    assert cnt == ref_cnt and np.array_equal(labels, ref_labels)


def test_group_500k_full_size_vs_oracle(ctx, orc):
    """BASELINE config 3 at full size: 500k planted hashes, threshold 31, vs the oracle's MIH search."""
    import os
    from rupphash_b200 import scanner
    n = 500_000
    hashes, low_conf = planted_hashes(n, seed=0xB200, n_clusters=5000, identical_block=1000)
    labels, cnt = scanner.group_labels(hashes, 31, low_conf=low_conf, ctx=ctx)
    ref_labels, ref_cnt, _ = orc.group_generic(hashes, 31, low_conf=low_conf, threads=min(32, os.cpu_count() or 1))
    assert cnt == ref_cnt
    assert np.array_equal(labels, ref_labels)
    # size-independent properties: labels are idempotent (label of a label is itself) and minimal
    assert np.array_equal(labels[labels], labels)
    assert np.all(labels <= np.arange(n))


def test_u64_group_and_find_groups_kats(ctx, orc):
    """hamminghash.rs:283-332 test_high_similarity_support, through the device path."""
    import ctypes as C
    from rupphash_b200 import _lib, hamminghash
    u = np.array([0, 0xFFF], np.uint64)
    groups = hamminghash.find_groups(hamminghash.MIHIndex.new(u), 12, ctx)
    assert len(groups) == 1 and sorted(groups[0]) == [0, 1]
    h = np.zeros((2, 32), np.uint8)
    h[1, 0:3] = 0xFF
    h[1, 3] = 0x3F  # 30 low bits
    groups = hamminghash.find_groups(hamminghash.MIHIndex.new(h), 30, ctx)
    assert len(groups) == 1 and sorted(groups[0]) == [0, 1]
    assert hamminghash.find_groups(hamminghash.MIHIndex.new(h), 29, ctx) == []
    # u64 grouping vs python brute force
    rng = np.random.default_rng(4)
    base = rng.integers(0, 2**63, size=400, dtype=np.uint64)
    base[100:150] = base[0:50] ^ np.uint64(0b1011)
    labels = np.empty(400, np.uint32)
    cnt = C.c_uint64()
    ctx.check(_lib.lib().rh_hamming_group_u64(ctx.handle, _lib.ptr(base), None, None, None, None, 400, 5,
                                              _lib.ptr(labels), C.byref(cnt)))
    par = list(range(400))
    def find(x):
        while par[x] != x:
            x = par[x]
        return x
    want_cnt = 0
    for i in range(400):
        for j in range(i + 1, 400):
            if bin(int(base[i]) ^ int(base[j])).count("1") <= 5:
                want_cnt += 1
                a, b = find(i), find(j)
                if a != b:
                    par[max(a, b)] = min(a, b)
    assert cnt.value == want_cnt
    assert labels.tolist() == [find(i) for i in range(400)]
    assert ctx.hamming_last_variant() == 1          # the one-POPC OR bound, chosen from the sampled selectivity
    try:                                            # ... and the exact-distance loop gives the same
        ctx.set_option("hamming.prefilter", 0)
        labels2 = np.empty(400, np.uint32)
        ctx.check(_lib.lib().rh_hamming_group_u64(ctx.handle, _lib.ptr(base), None, None, None, None, 400, 5,
                                                  _lib.ptr(labels2), C.byref(cnt)))
        assert cnt.value == want_cnt and np.array_equal(labels, labels2) and ctx.hamming_last_variant() == 0
    finally:
        ctx.set_option("hamming.prefilter", -1)
    # values that share their low 40 bits: the bound rejects nothing, the kernel falls back to the exact loop
    dense = (base & np.uint64(0xFFFFFF0000000000)) | np.uint64(0x123456789A)
    ctx.check(_lib.lib().rh_hamming_group_u64(ctx.handle, _lib.ptr(dense), None, None, None, None, 400, 15,
                                              _lib.ptr(labels2), C.byref(cnt)))
    assert ctx.hamming_last_variant() == 0
    d = np.array([[bin(int(a) ^ int(b)).count("1") for b in dense] for a in dense])
    assert cnt.value == int(((d <= 15).sum() - 400) // 2)


def test_find_groups_star_vs_oracle(ctx, orc):
    """find_groups is a greedy star clustering, not connected components (hamminghash.rs:245-270)."""
    from rupphash_b200 import hamminghash
    hashes, _ = planted_hashes(4000, seed=31, threshold=20)
    for max_dist in (8, 20, 31):
        got = hamminghash.find_groups(hamminghash.MIHIndex.new(hashes), max_dist, ctx)
        want = orc.MIHIndex(hashes).find_groups(max_dist, threads=4)
        assert [g[0] for g in got] == [g[0] for g in want]
        assert [sorted(g) for g in got] == [sorted(g) for g in want]
    u = np.random.default_rng(2).integers(0, 2**63, size=3000, dtype=np.uint64)
    u[1000:1400] = u[0:400] ^ np.uint64(0x8001)
    got = hamminghash.find_groups(hamminghash.MIHIndex.new(u), 5, ctx)
    want = orc.MIHIndex(u).find_groups(5, threads=2)
    assert [sorted(g) for g in got] == [sorted(g) for g in want]


def test_prefilter_variants_agree(ctx, orc):
    """The two-stage (prefix lower bound + refine) kernels and the full-distance kernel are the
    same function: identical labels, edge counts and edge multisets at the bench threshold."""
    from rupphash_b200 import scanner
    n = 30_000
    hashes, low_conf = planted_hashes(n, seed=41)
    hashes[5000:5200, 12:] = hashes[100, 12:]     # pairs that pass the 96-bit prefix test but fail the full one
    hashes[6000:6100, :12] = hashes[200, :12]     # prefix-identical, full distance large
    hashes[7000:7100, :24] = hashes[300, :24]     # the first 192 bits identical (OR bound 0), the rest random
    hashes[8000:8050, 24:] = hashes[400, 24:]     # only the last 64 bits shared
    try:
        for sim in (10, 31, 40, 63):
            ref_labels, ref_cnt, _ = orc.group_generic(hashes, sim, low_conf=low_conf, threads=4, use_mih=(sim <= 31))
            for pf in (0, 1, 2, 3, 4, 5, 6, 7):
                ctx.set_option("hamming.prefilter", pf)
                labels, cnt = scanner.group_labels(hashes, sim, low_conf=low_conf, ctx=ctx)
                assert cnt == ref_cnt, (sim, pf)
                assert np.array_equal(labels, ref_labels), (sim, pf)
                assert ctx.hamming_last_variant() == pf
    finally:
        ctx.set_option("hamming.prefilter", -1)


def test_adaptive_variant_on_unselective_prefix(ctx, orc):
    """Hashes that share their first 96 / 128 bits defeat the prefix filter; the kernel variant is chosen on
    the device from a sampled selectivity (full distance for such inputs) and the results stay exact."""
    from rupphash_b200 import scanner
    rng = np.random.default_rng(77)
    n = 20_000
    for shared in (12, 16):
        hashes = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        hashes[:, :shared] = hashes[0, :shared]
        hashes[100:140, shared:] = hashes[100, shared:]               # an identical block
        hashes[200:230, 31] ^= np.arange(30, dtype=np.uint8)          # near duplicates
        hashes[200:230, shared:31] = hashes[200, shared:31]
        ref_labels, ref_cnt, _ = orc.group_generic(hashes, 31, threads=4)
        labels, cnt = scanner.group_labels(hashes, 31, ctx=ctx)
        assert cnt == ref_cnt and np.array_equal(labels, ref_labels), shared
        # the prefix bounds reject nothing here, the OR bound over all 256 bits (3 POPC) still does:
        # 0 + 28 + 24 = 52 +- 3 bits (96 shared) / 0 + 24 + 24 = 48 +- 3.5 bits (128 shared) for unrelated pairs
        assert ctx.hamming_last_variant() == 7, (shared, ctx.hamming_last_variant())


def test_group_max_dist_matches_reference_rule(ctx, orc):
    """scanner.rs:2217-2241: max over members of min over the pivot's 8 variants (or of the plain
    pivot hash when it has no coefficients)."""
    from rupphash_b200 import scanner
    rng = np.random.default_rng(12)
    n = 400
    hashes, _ = planted_hashes(n, seed=3)
    coeffs = (rng.standard_normal((n, 256)) * 30).astype(np.float32)
    has_hash = (rng.random(n) > 0.1).astype(np.uint8)
    groups = [sorted(rng.choice(n, size=int(rng.integers(2, 9)), replace=False).tolist()) for _ in range(60)]
    pivots = [next(i for i in g if has_hash[i]) for g in groups]
    got = scanner.group_max_dist(groups, hashes, pivots, coefficients=coeffs, has_hash=has_hash, ctx=ctx)
    got_plain = scanner.group_max_dist(groups, hashes, pivots, has_hash=has_hash, ctx=ctx)
    for g, members in enumerate(groups):
        variants = orc.dihedral(coeffs[pivots[g]])
        want = max(min(orc.hamming256(v, hashes[i]) for v in variants) for i in members if has_hash[i])
        assert got[g] == want
        want_plain = max(orc.hamming256(hashes[pivots[g]], hashes[i]) for i in members if has_hash[i])
        assert got_plain[g] == want_plain
    assert len(scanner.group_max_dist([], hashes, [], ctx=ctx)) == 0


def test_group_500k_similarity_40_vs_mih_oracle(ctx, orc):
    """configs[2] size at the reference's DEFAULT similarity (40: MIH radius 2, the PF = 4 kernel)."""
    import os
    from rupphash_b200 import scanner
    n = 500_000
    hashes, low_conf = planted_hashes(n, seed=0xB200, n_clusters=5000, identical_block=1000, threshold=40)
    labels, cnt = scanner.group_labels(hashes, 40, low_conf=low_conf, ctx=ctx)
    ref_labels, ref_cnt, _ = orc.group_generic(hashes, 40, low_conf=low_conf, threads=os.cpu_count() or 4)
    assert cnt == ref_cnt
    assert np.array_equal(labels, ref_labels)


def test_group_150k_similarity_63_vs_bruteforce(ctx, orc):
    """The largest allowed similarity (63, the full-distance kernel) against the oracle's brute-force statement of
    the edge rule (the MIH probe count at radius 3 makes the CPU search slower than brute force at this size)."""
    import os
    from rupphash_b200 import scanner
    n = 150_000
    hashes, low_conf = planted_hashes(n, seed=63, threshold=63)
    labels, cnt = scanner.group_labels(hashes, 63, low_conf=low_conf, ctx=ctx)
    ref_labels, ref_cnt, _ = orc.group_generic(hashes, 63, low_conf=low_conf, threads=os.cpu_count() or 4, use_mih=False)
    assert cnt == ref_cnt
    assert np.array_equal(labels, ref_labels)


def test_group_500k_similarity_63_rows_and_closure(ctx, orc):
    """Full size at similarity 63 through size-independent properties: (1) for 1500 sampled files the device's edge
    list holds exactly the pairs (i, j > i) a CPU scan of that file's row of the pair matrix finds; (2) the labels
    are the connected components of the device's own edge list; (3) comparison_count = number of edges."""
    from rupphash_b200 import scanner
    n = 500_000
    hashes, low_conf = planted_hashes(n, seed=0x63, n_clusters=5000, identical_block=1000, threshold=63)
    labels, cnt = scanner.group_labels(hashes, 63, low_conf=low_conf, ctx=ctx)
    edges, cnt2 = scanner.edges(hashes, 63, low_conf=low_conf, cap=4_000_000, ctx=ctx)
    assert cnt == cnt2 == len(edges)
    assert (edges[:, 0] < edges[:, 1]).all()
    # (1) sampled rows: files inside planted structure and random ones
    rng = np.random.default_rng(5)
    sample = np.unique(np.concatenate([rng.integers(0, n, 700), edges[rng.integers(0, len(edges), 800), 0]]))
    h64 = hashes.view(np.uint64).reshape(n, 4)
    lc = low_conf.astype(bool)
    from_dev = {int(i): set() for i in sample}
    sel = np.isin(edges[:, 0], sample)
    for i, j in edges[sel]:
        from_dev[int(i)].add(int(j))
    for i in sample:
        x = h64[i + 1:] ^ h64[i]
        d = np.zeros(len(x), np.uint32)
        for w in range(4):
            v = x[:, w]
            v = v - ((v >> np.uint64(1)) & np.uint64(0x5555555555555555))
            v = (v & np.uint64(0x3333333333333333)) + ((v >> np.uint64(2)) & np.uint64(0x3333333333333333))
            v = (v + (v >> np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
            d += ((v * np.uint64(0x0101010101010101)) >> np.uint64(56)).astype(np.uint32)
        limit = np.where(lc[i + 1:] | lc[i], 0, 63)          # scanner.rs:1699, :1721
        want = set((np.flatnonzero(d <= limit) + i + 1).tolist())
        assert from_dev[int(i)] == want, int(i)
    # (2) labels = components of the edge list (min index of the component)
    parent = np.arange(n)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    for i, j in edges:
        ri, rj = find(int(i)), find(int(j))
        if ri != rj:
            parent[max(ri, rj)] = min(ri, rj)
    comp = np.array([find(i) for i in range(n)], np.uint32)
    assert np.array_equal(np.asarray(labels), comp)
