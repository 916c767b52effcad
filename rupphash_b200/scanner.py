"""Host mirror of the grouping core and feeder of src/scanner.rs over the CUDA library.

    quality_100 / is_low_confidence   scanner.rs:1416-1418, :1588-1594
    group_files_generic               scanner.rs:1640-1817 (edge phase + union-find -> groups)
    group_with_pdqhash                scanner.rs:1827-1832
    group_files_sharded               the same search tiled over the ranks of a process group
    hash_files_batched                the scanner.rs:1202-1521 hash loop restructured into batches

Groups are returned in the canonical form of SURVEY.md 8a: members ascending, groups ordered
by first member (the reference's HashMap order is nondeterministic).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import pdqhash
from ._lib import MAX_SIMILARITY_256, PDQ_MIN_QUALITY, default_context, is_torch_tensor, lib, ptr


def quality_100(q: float) -> int:
    """scanner.rs:1416-1418: (q * 100).round().clamp(0, 100) as u16 (round half away from zero)."""
    v = np.float32(q) * np.float32(100.0)
    r = math.floor(float(v) + 0.5) if v >= 0 else -math.floor(-float(v) + 0.5)
    return int(min(100, max(0, r)))


def is_low_confidence(quality100) -> bool:
    """scanner.rs:1588-1594 / :1631-1636: unknown quality counts as good."""
    return quality100 is not None and quality100 < PDQ_MIN_QUALITY


def labels_to_groups(labels) -> list:
    """labels[i] = smallest index of i's component -> groups with > 1 member (scanner.rs:1809-1817)."""
    labels = np.asarray(labels.cpu().numpy() if is_torch_tensor(labels) else labels).astype(np.int64)
    order = np.argsort(labels, kind="stable")
    sl = labels[order]
    cut = np.flatnonzero(np.diff(sl)) + 1
    groups = []
    for part in np.split(order, cut):
        if part.size > 1:
            groups.append(part.tolist())
    groups.sort(key=lambda g: g[0])
    return groups


def _u8(x):
    if x is None or is_torch_tensor(x):
        return x
    return np.ascontiguousarray(x, dtype=np.uint8)


def _labels_out(like, n):
    if is_torch_tensor(like) and like.is_cuda:
        import torch
        return torch.empty(n, dtype=torch.int32, device=like.device)
    return np.empty(n, np.uint32)


def group_labels(hashes, similarity, has_hash=None, variants=None, n_variants=None, low_conf=None, ctx=None):
    """-> (labels[n], comparison_count).  similarity > 63 raises ValueError (scanner.rs:1650-1655)."""
    ctx = ctx or default_context()
    hashes = _u8(hashes)
    n = int(hashes.shape[0])
    labels = _labels_out(hashes, n)
    cnt = C.c_uint64()
    ctx.check(lib().rh_hamming_group(ctx.handle, ptr(hashes), ptr(_u8(has_hash)), ptr(_u8(variants)),
                                     ptr(_u8(n_variants)), ptr(_u8(low_conf)), n, int(similarity), ptr(labels),
                                     C.byref(cnt)))
    return labels, int(cnt.value)


def group_files_generic(hashes, similarity, has_hash=None, variants=None, n_variants=None, low_conf=None, ctx=None):
    """scanner.rs:1640-1817 -> (groups, comparison_count)."""
    labels, cnt = group_labels(hashes, similarity, has_hash, variants, n_variants, low_conf, ctx)
    return labels_to_groups(labels), cnt


def group_with_pdqhash(hashes, similarity, coefficients=None, quality100=None, has_hash=None, ctx=None):
    """scanner.rs:1827-1832 with PdqStrategy (scanner.rs:1611-1637): files that carry cached
    coefficients query with their 8 dihedral variants, the others with their own hash; a file
    whose quality_100 < 50 is low-confidence."""
    ctx = ctx or default_context()
    variants = None
    if coefficients is not None:
        variants = pdqhash.dihedral_from_coeffs(coefficients, ctx)
    low_conf = None
    if quality100 is not None:
        q = np.asarray(quality100)
        low_conf = (q < PDQ_MIN_QUALITY).astype(np.uint8)
    return group_files_generic(hashes, similarity, has_hash=has_hash, variants=variants, low_conf=low_conf, ctx=ctx)


def regroup_from_cache(hash_values, coeff_values, q100, similarity, ctx=None):
    """Group a library from the reference's cache entries alone -- no pixels, no re-hashing.

    hash_values / coeff_values: per file, the plaintext hash_db / coeff_db value (cachefmt.py) or None;
    q100: per file, the TAG_DERIVED_PDQ_QUALITY short or None.  As in PdqStrategy (scanner.rs:1611-1637):
    a file with cached coefficients queries with its 8 dihedral variants (recomputed on the device from
    the coefficients), a file without them with its own hash; quality < 50 marks it low-confidence,
    unknown quality counts as good (scanner.rs:1631-1636)."""
    from . import cachefmt
    ctx = ctx or default_context()
    hashes, has_hash, coeffs, has_coeffs, q = cachefmt.load_cached(hash_values, coeff_values, q100)
    n = len(hashes)
    variants = np.zeros((n, 8, 32), np.uint8)
    variants[:, 0] = hashes
    n_variants = np.ones(n, np.uint8)
    idx = np.flatnonzero(has_coeffs & has_hash)
    if idx.size:
        variants[idx] = pdqhash.dihedral_from_coeffs(coeffs[idx], ctx)
        n_variants[idx] = 8
    low_conf = np.array([is_low_confidence(v) for v in q], np.uint8)
    return group_files_generic(hashes, similarity, has_hash=has_hash, variants=variants, n_variants=n_variants,
                               low_conf=low_conf, ctx=ctx)


def edges(hashes, similarity, has_hash=None, variants=None, n_variants=None, low_conf=None, cap=1 << 20, ctx=None):
    """Debug view: the (unordered) edge list of the same search -> (edges[k, 2], comparison_count)."""
    ctx = ctx or default_context()
    hashes = _u8(hashes)
    n = int(hashes.shape[0])
    out = np.zeros((cap, 2), np.uint32)
    cnt = C.c_uint64()
    ctx.check(lib().rh_hamming_edges(ctx.handle, ptr(hashes), ptr(_u8(has_hash)), ptr(_u8(variants)),
                                     ptr(_u8(n_variants)), ptr(_u8(low_conf)), n, int(similarity), ptr(out), cap,
                                     C.byref(cnt)))
    return out[: min(cap, cnt.value)], int(cnt.value)


def group_shard(hashes, similarity, rank, world, has_hash=None, variants=None, n_variants=None, low_conf=None,
                ctx=None):
    """One rank's tiles -> (local forest parent[n], local edge count)."""
    ctx = ctx or default_context()
    hashes = _u8(hashes)
    n = int(hashes.shape[0])
    parent = _labels_out(hashes, n)
    cnt = C.c_uint64()
    ctx.check(lib().rh_hamming_group_shard(ctx.handle, ptr(hashes), ptr(_u8(has_hash)), ptr(_u8(variants)),
                                           ptr(_u8(n_variants)), ptr(_u8(low_conf)), n, int(similarity), int(rank),
                                           int(world), ptr(parent), C.byref(cnt)))
    return parent, int(cnt.value)


def merge_forests(parents, ctx=None):
    """(world, n) forests -> canonical labels[n] (rh_uf_merge)."""
    ctx = ctx or default_context()
    world, n = int(parents.shape[0]), int(parents.shape[1])
    if not is_torch_tensor(parents):
        parents = np.ascontiguousarray(parents, dtype=np.uint32)
    labels = _labels_out(parents, n)
    ctx.check(lib().rh_uf_merge(ctx.handle, ptr(parents), world, n, ptr(labels)))
    return labels


def group_files_sharded(hashes, similarity, group=None, has_hash=None, variants=None, n_variants=None, low_conf=None,
                        ctx=None, shard_fn=None, merge_fn=None):
    """The search tiled over the ranks of a torch.distributed process group (one process per
    GPU): every rank scans its tiles, the n x u32 forests are all-gathered (NCCL over NVLink on
    GPUs, gloo in the CPU tests), each rank merges them and the edge counts are all-reduced.
    Labels are identical on every rank and to the single-GPU result.

    shard_fn / merge_fn let the CPU (gloo) tests substitute the oracle's rank kernel for the
    device calls; the product path never sets them.
    """
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    shard_fn = shard_fn or (lambda: group_shard(hashes, similarity, rank, world, has_hash, variants, n_variants,
                                                low_conf, ctx))
    parent, cnt = shard_fn()
    if is_torch_tensor(parent):
        local = parent
    else:
        local = torch.from_numpy(np.ascontiguousarray(parent).view(np.int32))
    local = local.contiguous()
    gathered = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(gathered, local, group=group)   # NCCL: one n x u32 block per rank
    else:
        dist.all_gather(list(gathered.unbind(0)), local, group=group)
    total = torch.tensor([cnt], dtype=torch.int64, device=local.device)
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    if merge_fn is not None:
        labels = merge_fn(gathered.cpu().numpy().view(np.uint32))
    elif gathered.is_cuda:
        labels = merge_forests(gathered, ctx)
    else:
        labels = merge_forests(gathered.numpy().view(np.uint32), ctx)
    return labels, int(total.item())


def group_labels_multi(group, hashes, similarity, has_hash=None, variants=None, n_variants=None, low_conf=None,
                       out=None):
    """group_labels over every GPU of an rh_group (`_lib.Group`): one process, in-library NCCL over
    NVLink, tiles claimed by the GPUs from one pool (rh_hamming_group_multi).  Inputs: numpy arrays
    (ideally page-locked) or CUDA tensors on one GPU of the group.  -> (labels[n], comparison_count),
    bit-identical to the single-GPU result."""
    hashes = _u8(hashes)
    n = int(hashes.shape[0])
    labels = out if out is not None else np.empty(n, np.uint32)
    cnt = C.c_uint64()
    group.check(lib().rh_hamming_group_multi(group.handle, ptr(hashes), ptr(_u8(has_hash)), ptr(_u8(variants)),
                                             ptr(_u8(n_variants)), ptr(_u8(low_conf)), n, int(similarity),
                                             ptr(labels), C.byref(cnt)))
    return labels, int(cnt.value)


def merge_groups_by_stem(groups, paths):
    """scanner.rs:1905-1983: groups that contain files with the same parent directory AND the same
    file stem (e.g. IMG_1.jpg / IMG_1.cr2) are merged.  Host logic on paths; returns the groups in
    canonical form (members ascending and de-duplicated, groups ordered by first member)."""
    import os
    if len(groups) < 2:
        return [sorted(set(g)) for g in groups]
    parent = list(range(len(groups)))

    def find(i):
        root = i
        while parent[root] != root:
            root = parent[root]
        while parent[i] != root:
            parent[i], i = root, parent[i]
        return root

    first_group_of = {}
    for g_idx, group in enumerate(groups):
        for f_idx in group:
            p = paths[f_idx]
            key = (os.path.dirname(p), os.path.splitext(os.path.basename(p))[0])
            if not key[1]:
                continue
            other = first_group_of.setdefault(key, g_idx)
            if other != g_idx:
                ri, rj = find(other), find(g_idx)
                if ri != rj:
                    parent[ri] = rj
    merged = {}
    for g_idx, group in enumerate(groups):
        merged.setdefault(find(g_idx), []).extend(group)
    out = [sorted(set(g)) for g in merged.values()]
    out.sort(key=lambda g: g[0])
    return out


def group_max_dist(groups, hashes, pivots, coefficients=None, has_hash=None, ctx=None):
    """max_dist of every group (scanner.rs:2217-2241): max over the group's hashed members of the
    minimum distance to the 8 dihedral variants of the group's pivot (or to the pivot's plain
    hash when it has no cached coefficients).

    groups: list of index lists; pivots: one file index per group (the caller's sort picks it:
    the first member with features, else the first member with a hash); coefficients: (n, 256) f32
    or None."""
    ctx = ctx or default_context()
    hashes = _u8(hashes).reshape(-1, 32)
    ng = len(groups)
    if ng == 0:
        return np.zeros(0, np.uint32)
    pivots = np.asarray(pivots, np.int64)
    piv = np.zeros((ng, 8, 32), np.uint8)
    piv[:, 0, :] = hashes[pivots]
    nv = np.ones(ng, np.uint8)
    if coefficients is not None:
        piv[:] = np.asarray(pdqhash.dihedral_from_coeffs(np.ascontiguousarray(coefficients[pivots], np.float32), ctx))
        nv[:] = 8
    mem, grp = [], []
    for g, members in enumerate(groups):
        for i in members:
            if has_hash is None or has_hash[i]:
                mem.append(i)
                grp.append(g)
    mem_h = np.ascontiguousarray(hashes[np.asarray(mem, np.int64)]) if mem else np.zeros((0, 32), np.uint8)
    grp = np.asarray(grp, np.uint32)
    out = np.zeros(ng, np.uint32)
    ctx.check(lib().rh_group_max_dist(ctx.handle, ptr(piv), ptr(nv), ptr(mem_h), ptr(grp), len(mem), ng, ptr(out)))
    return out


def hash_files_batched(images_iter, batch_size=256, want_coeffs=True, ctx=None, progress=None, depth=2, hasher=None,
                       pinned=None):
    """Scanner-style feeder (scanner.rs:1202-1521 restructured): decoded images of mixed sizes
    arrive one by one (the decode stays on the host, as in the reference); same-sized images
    are collected into batches of `batch_size`, hashed on the device, and handed back in arrival
    order as dicts(hash, quality, quality_100, coeffs) -- or None where the reference returns
    None (scanner.rs:1481-1487 keeps the file without a hash).

    The reference overlaps decode and hashing with a rayon pool feeding a DbUpdate channel
    (scanner.rs:1202-1211, :1495-1518); here the caller's iterator (the decode) runs on the calling
    thread while ONE submitter thread owns the rh_ctx and hashes full batches, `depth` batches may be
    queued between them, and batches are staged in recycled page-locked buffers so that the library's
    H2D copies run at the pinned rate.  Progress ticks (scanner.rs:1206-1211) fire per finished batch as
    progress(done, seen).  `hasher(batch_array) -> hash_batch-style dict` replaces the device (tests);
    `pinned` defaults to True with the device hasher."""
    import queue
    import threading
    from ._lib import pinned_empty, pinned_free
    if hasher is None:
        ctx = ctx or default_context()
        hasher = lambda arr: pdqhash.hash_batch(arr, want_coeffs=want_coeffs, ctx=ctx)   # noqa: E731
        pinned = True if pinned is None else pinned
    pinned = bool(pinned)
    results = {}
    seen = [0]
    lock = threading.Lock()
    work = queue.Queue(maxsize=max(1, depth))
    free_bufs = {}            # nbytes -> [arrays]: recycled staging buffers
    failure = []

    def take_buffer(shape):
        nbytes = int(np.prod(shape))
        with lock:
            pool = free_bufs.get(nbytes)
            if pool:
                return pool.pop().reshape(shape)
        return pinned_empty(shape, np.uint8) if pinned else np.empty(shape, np.uint8)

    def submitter():
        while True:
            item = work.get()
            if item is None:
                return
            idxs, buf = item
            if failure:
                continue      # drain the queue after an error so that the producer never blocks
            try:
                out = hasher(buf[: len(idxs)])
                valid, quality, hashes = np.asarray(out["valid"]), np.asarray(out["quality"]), np.asarray(out["hash"])
                coeffs = None if out.get("coeffs") is None else np.asarray(out["coeffs"])
                with lock:
                    for k, i in enumerate(idxs):
                        if not valid[k]:
                            results[i] = None
                            continue
                        q = float(quality[k])
                        results[i] = {"hash": hashes[k].copy(), "quality": q, "quality_100": quality_100(q),
                                      "coeffs": coeffs[k].copy() if coeffs is not None else None}
                    free_bufs.setdefault(buf.size, []).append(buf.reshape(-1))
                    done, total = len(results), seen[0]
                if progress:
                    progress(done, total)
            except BaseException as e:   # handed to the caller
                failure.append(e)

    th = threading.Thread(target=submitter, name="rh-submitter", daemon=True)
    th.start()
    pending = {}              # image shape -> (indices, staging buffer)
    try:
        for i, img in enumerate(images_iter):
            if failure:
                break
            with lock:
                seen[0] = i + 1
            img = np.asarray(img, dtype=np.uint8)
            h, w = img.shape[:2]
            if w < pdqhash.MIN_HASHABLE_DIM or h < pdqhash.MIN_HASHABLE_DIM:
                with lock:
                    results[i] = None
                continue
            key = img.shape
            slot = pending.get(key)
            if slot is None:
                slot = pending[key] = ([], take_buffer((batch_size,) + key))
            slot[1][len(slot[0])] = img          # the copy into the staging buffer
            slot[0].append(i)
            if len(slot[0]) >= batch_size:
                work.put(pending.pop(key))
        for key in list(pending):
            work.put(pending.pop(key))
    finally:
        work.put(None)
        th.join()
        if pinned:
            for pool in free_bufs.values():
                for b in pool:
                    pinned_free(b)
    if failure:
        raise failure[0]
    return [results[i] for i in range(seen[0])]
