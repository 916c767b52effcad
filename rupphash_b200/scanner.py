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


class _PinnedPool:
    """Every page-locked buffer the feeder allocates, so that all of them are released on every exit path."""

    def __init__(self, pinned: bool):
        self.pinned = pinned
        self._all = []

    def empty(self, shape, dtype=np.uint8):
        from ._lib import pinned_empty
        if not self.pinned:
            return np.empty(shape, dtype)
        arr = pinned_empty(shape, dtype)
        self._all.append(arr)
        return arr

    def close(self):
        from ._lib import pinned_free
        for arr in self._all:
            pinned_free(arr)
        self._all = []


class _Slot:
    """One open batch: a staging buffer of `cap` same-shaped images being filled by the decode workers."""
    __slots__ = ("key", "buf", "cap", "idxs", "filled", "stamp")

    def __init__(self, key, buf, cap):
        self.key, self.buf, self.cap = key, buf, cap
        self.idxs, self.filled, self.stamp = [], 0, 0


class _DeviceHasher:
    """rh_pdq_hash_batch_async with results in recycled page-locked arrays: submit() queues a batch and
    returns at once, wait() blocks until that batch's results are in host memory."""

    def __init__(self, ctx, want_coeffs, pool, max_images):
        self.ctx, self.want_coeffs, self.pool, self.max_images = ctx, want_coeffs, pool, max_images
        self.free = []

    def _outs(self):
        if self.free:
            return self.free.pop()
        m = self.max_images
        return {"hash": self.pool.empty((m, 32)), "quality": self.pool.empty((m,), np.float32),
                "valid": self.pool.empty((m,)),
                "coeffs": self.pool.empty((m, 256), np.float32) if self.want_coeffs else None}

    def submit(self, arr):
        n, h, w = arr.shape[0], arr.shape[1], arr.shape[2]
        o = self._outs()
        ticket = C.c_uint64()
        self.ctx.check(lib().rh_pdq_hash_batch_async(
            self.ctx.handle, ptr(arr), pdqhash._layout_of(arr.shape[1:]), n, w, h, 0, 0, ptr(o["hash"]), ptr(o["quality"]),
            ptr(o["coeffs"]), None, ptr(o["valid"]), C.byref(ticket)))
        return ticket.value, o, n

    def wait(self, handle):
        ticket, o, n = handle
        self.ctx.check(lib().rh_ctx_wait(self.ctx.handle, ticket))
        res = {k: (None if v is None else v[:n].copy()) for k, v in o.items()}
        self.free.append(o)
        return res


class Feeder:
    """Keeps the feeder's page-locked staging and result buffers between hash_files_batched calls (page-locking a
    256 MB batch costs ~0.1 s: a scan that runs in several calls, or a benchmark's timed repeat, should not pay
    it again).  close() releases everything."""

    def __init__(self, pinned: bool = True):
        self.pool = _PinnedPool(bool(pinned))
        self.free_bufs = {}
        self.device_hasher = None

    def close(self):
        self.pool.close()
        self.free_bufs = {}
        self.device_hasher = None


def hash_files_batched(items, batch_size=256, want_coeffs=True, ctx=None, progress=None, depth=2, hasher=None,
                       pinned=None, decode=None, workers=0, batch_bytes=256 << 20, max_open_shapes=8, inflight=2,
                       feeder=None):
    """Scanner-style feeder (scanner.rs:1202-1521 restructured): files are decoded on the host (as in the
    reference), same-sized images are collected into batches, hashed on the device, and handed back in
    arrival order as dicts(hash, quality, quality_100, coeffs) -- or None where the reference returns None
    (scanner.rs:1481-1487 keeps the file without a hash).

    items    an iterable.  With decode=None it yields decoded images and is consumed on the calling thread.
             With `decode` (item -> HxW[xC] uint8 array or None) and `workers` > 0 a pool of decode threads
             pulls items and decodes them in parallel -- the reference's rayon pool (scanner.rs:1188-1205) --
             and copies each image straight into a slot of the page-locked staging batch of its shape.
    batches  at most `batch_size` images AND at most `batch_bytes` of pixels each (large photos get small
             batches), at most `max_open_shapes` partly filled batches at a time (the least recently used one
             is sent early), staging buffers are recycled, and every page-locked buffer is released on exit.
    device   ONE submitter thread owns the rh_ctx; it keeps `inflight` batches queued with
             rh_pdq_hash_batch_async so that the copy of batch k+1 runs under the kernels of batch k.
    progress ticks (scanner.rs:1206-1211) fire per finished batch as progress(done, seen).
    `hasher(batch_array) -> hash_batch-style dict` replaces the device (CPU tests); `pinned` defaults to True
    with the device.  `feeder` (a Feeder) keeps the page-locked buffers for the next call."""
    import queue
    import threading
    device = hasher is None
    if device:
        ctx = ctx or default_context()
        pinned = True if pinned is None else pinned
    own_feeder = feeder is None
    if feeder is None:
        feeder = Feeder(bool(pinned))
    pool = feeder.pool
    results, failure = {}, []
    seen = [0]
    lock = threading.Condition()
    work = queue.Queue(maxsize=max(1, depth))
    free_bufs = feeder.free_bufs   # (cap, shape) -> [staging arrays]
    open_slots = {}           # shape -> _Slot
    stamp = [0]

    def publish(idxs, out):
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
            done, total = len(results), seen[0]
        if progress:
            progress(done, total)

    def recycle(slot):
        with lock:
            free_bufs.setdefault((slot.cap, slot.key), []).append(slot.buf)

    def submitter():
        dev = None
        pending = []          # queued device batches, oldest first
        try:
            while True:
                item = work.get()
                if item is None:
                    break
                if failure:
                    continue      # drain the queue after an error so that the producers never block
                slot = item
                arr = slot.buf[: len(slot.idxs)]
                if device:
                    if dev is None:
                        cap = max(1, min(batch_size, 1 << 16))
                        dev = feeder.device_hasher
                        if dev is None or dev.max_images < cap or dev.want_coeffs != want_coeffs or dev.ctx is not ctx:
                            dev = feeder.device_hasher = _DeviceHasher(ctx, want_coeffs, pool, cap)
                    pending.append((dev.submit(arr), slot))
                    if len(pending) >= max(1, inflight):
                        h, s = pending.pop(0)
                        publish(s.idxs, dev.wait(h))
                        recycle(s)
                else:
                    publish(slot.idxs, hasher(arr))
                    recycle(slot)
            while pending and not failure:
                h, s = pending.pop(0)
                publish(s.idxs, dev.wait(h))
                recycle(s)
        except BaseException as e:   # handed to the caller
            failure.append(e)
            while True:               # keep draining until the sentinel so that producers never block
                try:
                    if work.get(timeout=0.05) is None:
                        break
                except queue.Empty:
                    if done_feeding[0]:
                        break
        finally:
            if device and pending:
                try:
                    ctx.sync()
                except Exception:
                    pass

    done_feeding = [False]

    def send(slot):
        """slot has left open_slots: wait for the copies in progress, then queue it"""
        with lock:
            while slot.filled < len(slot.idxs):
                lock.wait()
        if slot.idxs:
            work.put(slot)

    def stage(i, img):
        """copy one decoded image into the open batch of its shape; full / evicted batches go to the submitter"""
        if img is None:
            with lock:
                results[i] = None
            return
        img = np.asarray(img, dtype=np.uint8)
        h, w = img.shape[:2]
        if w < pdqhash.MIN_HASHABLE_DIM or h < pdqhash.MIN_HASHABLE_DIM:
            with lock:
                results[i] = None
            return
        key = img.shape
        evicted = None
        with lock:
            slot = open_slots.get(key)
            if slot is None:
                if len(open_slots) >= max(1, max_open_shapes):      # send the least recently used batch early
                    lru = min(open_slots.values(), key=lambda s: s.stamp)
                    evicted = open_slots.pop(lru.key)
                cap = int(max(1, min(batch_size, batch_bytes // max(1, img.nbytes))))
                bufs = free_bufs.get((cap, key))
                buf = bufs.pop() if bufs else None
                slot = open_slots[key] = _Slot(key, buf, cap)
            k = len(slot.idxs)
            slot.idxs.append(i)
            stamp[0] += 1
            slot.stamp = stamp[0]
            full = len(slot.idxs) >= slot.cap
            if full:
                del open_slots[key]
            creator = slot.buf is None and k == 0
        if creator:                                                 # a (blocking) page-locked allocation: not under the lock
            buf = pool.empty((slot.cap,) + key, np.uint8)
            with lock:
                slot.buf = buf
                lock.notify_all()
        else:
            with lock:
                while slot.buf is None:
                    lock.wait()
        slot.buf[k] = img                                           # the copy into the staging buffer
        with lock:
            slot.filled += 1
            lock.notify_all()
        if evicted is not None:
            send(evicted)
        if full:
            send(slot)

    th = threading.Thread(target=submitter, name="rh-submitter", daemon=True)
    th.start()
    decoders = []
    try:
        if decode is not None and workers and workers > 0:
            it = iter(items)
            it_lock = threading.Lock()

            def decoder():
                while not failure:
                    with it_lock:
                        try:
                            item = next(it)
                        except StopIteration:
                            return
                        except BaseException as e:
                            failure.append(e)
                            return
                        with lock:
                            i = seen[0]
                            seen[0] = i + 1
                    try:
                        stage(i, decode(item))
                    except BaseException as e:
                        failure.append(e)
                        return
            decoders = [threading.Thread(target=decoder, name=f"rh-decode-{k}", daemon=True) for k in range(int(workers))]
            for d in decoders:
                d.start()
            for d in decoders:
                d.join()
        else:
            for i, item in enumerate(items):
                if failure:
                    break
                with lock:
                    seen[0] = i + 1
                stage(i, decode(item) if decode is not None else item)
        if not failure:
            with lock:
                rest = sorted(open_slots.values(), key=lambda s: s.stamp)
                open_slots.clear()
            for slot in rest:
                send(slot)
    finally:
        done_feeding[0] = True
        work.put(None)
        th.join()
        if own_feeder:
            feeder.close()
    if failure:
        raise failure[0]
    return [results[i] for i in range(seen[0])]


# ------------------------------------------------------------------ group post-processing ----
# scanner.rs:1986-2022 process_raw_groups, :2183-2254 analyze_group_with_features, :2040-2110 sort_files,
# :2256-2262 sort_by_stem_then_ext, :1561-1576 the final group order.  Path and metadata logic runs on the
# host (it is string work); the one data-parallel piece, max_dist over every group, is a single
# rh_group_max_dist call for the whole library.

RAW_EXTS = ("nef", "dng", "cr2", "cr3", "arw", "orf", "rw2", "raf", "kdc", "dcr", "pef", "x3f", "srf", "3fr")  # scanner.rs:43-46
STATUS_NONE, STATUS_SOME_IDENTICAL, STATUS_ALL_IDENTICAL = "None", "SomeIdentical", "AllIdentical"


class FileMeta:
    """The fields of FileMetadata (scanner.rs to_file_metadata) that grouping output depends on."""
    __slots__ = ("path", "size", "modified", "content_hash", "pixel_hash", "pdqhash", "exif_timestamp", "index")

    def __init__(self, path, size=0, modified=0, content_hash=b"", pixel_hash=None, pdqhash=None, exif_timestamp=None,
                 index=None):
        self.path, self.size, self.modified = path, size, modified
        self.content_hash, self.pixel_hash, self.pdqhash = content_hash, pixel_hash, pdqhash
        self.exif_timestamp, self.index = exif_timestamp, index


def _file_name(path: str) -> str:
    import os
    return os.path.basename(path)


def _file_stem(path: str) -> str:
    import os
    return os.path.splitext(os.path.basename(path))[0]


def is_raw_ext(path: str) -> bool:
    """scanner.rs:2264-2269"""
    import os
    ext = os.path.splitext(os.path.basename(path))[1]
    return ext[1:].lower() in RAW_EXTS if ext else False


def natural_key(s: str):
    """Sort key with the order of natord::compare (Martin Pool's strnatcmp, case-sensitive, whitespace
    skipped): digit runs compare as numbers -- right-aligned (longer run is larger, then the first differing
    digit) unless one of them starts with '0', in which case they compare left-aligned like fractions.
    The `natord` crate is not in the reference tree; this restates its published algorithm."""
    import functools

    def cmp(a: str, b: str) -> int:
        ai = bi = 0
        na, nb = len(a), len(b)
        while True:
            while ai < na and a[ai].isspace():
                ai += 1
            while bi < nb and b[bi].isspace():
                bi += 1
            ca = a[ai] if ai < na else ""
            cb = b[bi] if bi < nb else ""
            if ca.isdigit() and cb.isdigit():
                if ca == "0" or cb == "0":        # left-aligned ("fractional") comparison
                    while True:
                        da = a[ai] if ai < na else ""
                        db = b[bi] if bi < nb else ""
                        if not da.isdigit() and not db.isdigit():
                            r = 0
                            break
                        if not da.isdigit():
                            return -1
                        if not db.isdigit():
                            return 1
                        if da != db:
                            return -1 if da < db else 1
                        ai += 1
                        bi += 1
                else:                              # right-aligned: the longer run wins, else the first difference
                    bias = 0
                    while True:
                        da = a[ai] if ai < na else ""
                        db = b[bi] if bi < nb else ""
                        if not da.isdigit() and not db.isdigit():
                            r = bias
                            break
                        if not da.isdigit():
                            return -1
                        if not db.isdigit():
                            return 1
                        if bias == 0 and da != db:
                            bias = -1 if da < db else 1
                        ai += 1
                        bi += 1
                if r:
                    return r
                continue
            if not ca and not cb:
                return 0
            if ca != cb:
                return -1 if ca < cb else 1
            ai += 1
            bi += 1
    return functools.cmp_to_key(cmp)(s)


def sort_files(files: list, sort_order: str) -> None:
    """scanner.rs:2040-2110, in place; every branch is a stable sort as in the reference ("random" and
    "location" leave the order alone: the reference shuffles / defers to the GUI)."""
    name = lambda f: _file_name(f.path)                    # noqa: E731
    nat = lambda f: natural_key(_file_name(f.path))        # noqa: E731
    if sort_order == "name":
        files.sort(key=name)
    elif sort_order == "name-desc":
        files.sort(key=name)
        files.reverse()
    elif sort_order == "name-natural":
        files.sort(key=nat)
    elif sort_order == "name-natural-desc":
        files.sort(key=nat)
        files.reverse()
    elif sort_order == "date":
        files.sort(key=lambda f: f.modified)
    elif sort_order == "date-desc":
        files.sort(key=lambda f: f.modified, reverse=True)   # sort_by_key(Reverse(..)): stable, ties keep their order
    elif sort_order == "size":
        files.sort(key=lambda f: f.size)
    elif sort_order == "size-desc":
        files.sort(key=lambda f: f.size, reverse=True)
    elif sort_order in ("exif-date", "exif-date-desc"):
        sign = 1 if sort_order == "exif-date" else -1
        files.sort(key=lambda f: (0, sign * f.exif_timestamp) if f.exif_timestamp is not None else (1, sign * f.modified))
    elif sort_order in ("random", "location"):
        pass
    else:
        files.sort(key=nat)


def sort_by_stem_then_ext(files: list) -> None:
    """scanner.rs:2256-2262: same stem together, the RAW file after its JPEG."""
    files.sort(key=lambda f: (_file_stem(f.path), is_raw_ext(f.path)))


def _arrange_group(files: list, sort_order: str):
    """The ordering half of analyze_group_with_features (scanner.rs:2189-2212) -> (files, status)."""
    counts = {}
    for f in files:
        counts[f.content_hash] = counts.get(f.content_hash, 0) + 1
    duplicates = [f for f in files if counts[f.content_hash] > 1]
    unique = [f for f in files if counts[f.content_hash] <= 1]
    duplicates.sort(key=lambda f: ((0, b"") if f.pixel_hash is None else (1, bytes(f.pixel_hash)), bytes(f.content_hash),
                                   _file_name(f.path)))
    sort_files(unique, sort_order)
    out = duplicates + unique
    sort_by_stem_then_ext(out)
    if len(counts) == 1:
        status = STATUS_ALL_IDENTICAL
    elif any(c > 1 for c in counts.values()):
        status = STATUS_SOME_IDENTICAL
    else:
        status = STATUS_NONE
    return out, status


def process_raw_groups(raw_groups, files, sort_order="name-natural", coefficients=None, ctx=None, max_dist_fn=None,
                       dihedral_fn=None):
    """scanner.rs:1986-2022 -> (groups of FileMeta in display order, infos [{max_dist, status}]).

    raw_groups: index lists (the output of group_files_generic / merge_groups_by_stem); files: FileMeta per
    scanned file; coefficients: {file index: 256 f32} for the files that carry cached PDQ features
    (features_map, scanner.rs:1995-2000).  Every group is ordered like analyze_group_with_features does, its
    pivot is the first file with features (8 dihedral variants) or else the first file with a hash
    (scanner.rs:2217-2241), and max_dist of ALL groups comes from one rh_group_max_dist call.
    `max_dist_fn(pivot_variants, n_variants, member_hashes, member_group, n_groups)` and `dihedral_fn(coeffs)`
    replace the device calls (CPU tests of the host logic); the product path never sets them."""
    coefficients = coefficients or {}
    groups, statuses = [], []
    for idxs in raw_groups:
        members = []
        for i in idxs:
            f = files[i]
            if f.index is None:
                f.index = i
            members.append(f)
        arranged, status = _arrange_group(members, sort_order.lower())
        groups.append(arranged)
        statuses.append(status)
    ng = len(groups)
    piv = np.zeros((ng, 8, 32), np.uint8)
    nv = np.ones(ng, np.uint8)
    feat_groups, feat_coeffs = [], []
    mem_h, mem_g = [], []
    for g, arranged in enumerate(groups):
        with_feat = next((f for f in arranged if f.index in coefficients), None)
        if with_feat is not None:
            feat_groups.append(g)
            feat_coeffs.append(np.asarray(coefficients[with_feat.index], np.float32).reshape(256))
        else:
            with_hash = next((f for f in arranged if f.pdqhash is not None), None)
            if with_hash is not None:
                piv[g, 0] = np.frombuffer(bytes(with_hash.pdqhash), np.uint8)
            else:
                nv[g] = 0                     # no pivot: max_dist 0 (scanner.rs:2239-2241)
        for f in arranged:
            if f.pdqhash is not None:
                mem_h.append(np.frombuffer(bytes(f.pdqhash), np.uint8))
                mem_g.append(g)
    if feat_groups:
        if dihedral_fn is not None:
            var = dihedral_fn(np.stack(feat_coeffs))
        else:
            ctx = ctx or default_context()
            var = pdqhash.dihedral_from_coeffs(np.stack(feat_coeffs), ctx)
        piv[np.asarray(feat_groups)] = np.asarray(var)
        nv[np.asarray(feat_groups)] = 8
    mem_h = np.stack(mem_h) if mem_h else np.zeros((0, 32), np.uint8)
    mem_g = np.asarray(mem_g, np.uint32)
    # groups without a pivot take no part in the reduce
    keep = nv[mem_g] > 0 if len(mem_g) else np.zeros(0, bool)
    mem_h, mem_g = np.ascontiguousarray(mem_h[keep]), np.ascontiguousarray(mem_g[keep])
    nv_call = np.maximum(nv, 1).astype(np.uint8)
    if ng == 0:
        max_dist = np.zeros(0, np.uint32)
    elif max_dist_fn is not None:
        max_dist = np.asarray(max_dist_fn(piv, nv_call, mem_h, mem_g, ng), np.uint32)
    else:
        ctx = ctx or default_context()
        max_dist = np.zeros(ng, np.uint32)
        ctx.check(lib().rh_group_max_dist(ctx.handle, ptr(piv), ptr(nv_call), ptr(mem_h), ptr(mem_g), len(mem_g), ng,
                                          ptr(max_dist)))
    infos = [{"max_dist": int(max_dist[g]) if nv[g] else 0, "status": statuses[g]} for g in range(ng)]
    return groups, infos


def sort_groups(groups, infos):
    """The final order of scan_and_group (scanner.rs:1561-1576): groups with identical files first, then
    by max_dist ascending, then by the size of the first file descending (stable)."""
    order = sorted(range(len(groups)), key=lambda g: (0 if infos[g]["status"] != STATUS_NONE else 1, infos[g]["max_dist"],
                                                       -(groups[g][0].size if groups[g] else 0)))
    return [groups[g] for g in order], [infos[g] for g in order]


def scan_groups(files, similarity, sort_order="name-natural", coefficients=None, quality100=None, ctx=None):
    """The grouping half of scan_and_group (scanner.rs:1542-1576) on already hashed files: edge phase +
    union-find on the device, stem merge, per-group ordering, max_dist, final group order -- the list
    `phdupes` prints.  files: FileMeta per scanned file (pdqhash None = not hashed)."""
    ctx = ctx or default_context()
    n = len(files)
    hashes = np.zeros((n, 32), np.uint8)
    has_hash = np.zeros(n, np.uint8)
    for i, f in enumerate(files):
        f.index = i
        if f.pdqhash is not None:
            hashes[i] = np.frombuffer(bytes(f.pdqhash), np.uint8)
            has_hash[i] = 1
    coefficients = coefficients or {}
    variants = np.zeros((n, 8, 32), np.uint8)
    variants[:, 0] = hashes
    n_variants = np.ones(n, np.uint8)
    idx = [i for i in sorted(coefficients) if has_hash[i]]
    if idx:
        variants[np.asarray(idx)] = pdqhash.dihedral_from_coeffs(
            np.stack([np.asarray(coefficients[i], np.float32).reshape(256) for i in idx]), ctx)
        n_variants[np.asarray(idx)] = 8
    low_conf = None
    if quality100 is not None:
        low_conf = np.array([1 if is_low_confidence(q) else 0 for q in quality100], np.uint8)
    raw, comparisons = group_files_generic(hashes, similarity, has_hash=has_hash, variants=variants, n_variants=n_variants,
                                           low_conf=low_conf, ctx=ctx)
    raw = merge_groups_by_stem(raw, [f.path for f in files])
    groups, infos = process_raw_groups(raw, files, sort_order, coefficients, ctx)
    groups, infos = sort_groups(groups, infos)
    return groups, infos, comparisons
