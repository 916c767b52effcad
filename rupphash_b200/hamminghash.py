"""Host mirror of src/hamminghash.rs over the CUDA library.

    hamming_distance / hamming_distances    HammingHash::hamming_distance  hamminghash.rs:34-36, :55-58
    get_chunk                               HammingHash::get_chunk         hamminghash.rs:29-32, :50-53
    MIHIndex                                hamminghash.rs:82-149 (new / bucket / hash / len)
    find_groups                             hamminghash.rs:191-271

On the device the pair search is an exact all-pairs scan, so MIHIndex only has to hold the
hashes; bucket() is kept for API parity and built lazily on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import MAX_SIMILARITY_64, MAX_SIMILARITY_256, default_context, is_torch_tensor, lib, ptr

__all__ = ["MAX_SIMILARITY_64", "MAX_SIMILARITY_256", "hamming_distance", "hamming_distances", "get_chunk",
           "MIHIndex", "find_groups"]


def _as_hashes(h):
    """-> (array, width_bits).  uint64 arrays are 64-bit hashes, (n, 32) uint8 are PDQ hashes."""
    if is_torch_tensor(h):
        import torch
        if h.dtype in (torch.int64, torch.uint64):
            return h.contiguous(), 64
        return h.contiguous().view(-1, 32), 256
    h = np.asarray(h)
    if h.dtype == np.uint64:
        return np.ascontiguousarray(h).reshape(-1), 64
    return np.ascontiguousarray(h, dtype=np.uint8).reshape(-1, 32), 256


def hamming_distances(a, b, ctx=None):
    """Element-wise popcount(a[i] ^ b[i]) -> uint32[n]."""
    ctx = ctx or default_context()
    a, wa = _as_hashes(a)
    b, wb = _as_hashes(b)
    if wa != wb or a.shape[0] != b.shape[0]:
        raise ValueError("hash arrays differ in width or length")
    n = a.shape[0]
    if is_torch_tensor(a) and a.is_cuda:
        import torch
        out = torch.empty(n, dtype=torch.int32, device=a.device)
    else:
        out = np.empty(n, np.uint32)
    fn = lib().rh_hamming_distances if wa == 256 else lib().rh_hamming_distances_u64
    ctx.check(fn(ctx.handle, ptr(a), ptr(b), n, ptr(out)))
    return out


def hamming_distance(a, b, ctx=None) -> int:
    """One pair (batch of 1 on the device)."""
    if isinstance(a, (int, np.integer)):
        a, b = np.array([a], np.uint64), np.array([b], np.uint64)
    else:
        a, b = np.asarray(a, np.uint8).reshape(1, 32), np.asarray(b, np.uint8).reshape(1, 32)
    return int(hamming_distances(a, b, ctx)[0])


def get_chunk(h, k: int) -> int:
    """hamminghash.rs:29-32 (u64: byte k) / :50-53 ([u8;32]: u16 little-endian at bytes 2k, 2k+1)."""
    if isinstance(h, (int, np.integer)):
        return (int(h) >> (8 * k)) & 0xFF
    h = np.asarray(h, np.uint8)
    return int(h[2 * k]) | (int(h[2 * k + 1]) << 8)


class MIHIndex:
    """hamminghash.rs:82-149."""

    def __init__(self, hashes):
        self.hashes, self.width = _as_hashes(hashes)
        self.num_chunks = 8 if self.width == 64 else 16
        self.num_buckets = 256 if self.width == 64 else 65536
        self._csr = None

    @classmethod
    def new(cls, hashes):
        return cls(hashes)

    def __len__(self):
        return int(self.hashes.shape[0])

    def len(self):
        return len(self)

    def hash(self, dense_id: int):
        return self.hashes[dense_id]

    def _chunks(self) -> np.ndarray:
        h = self.hashes.cpu().numpy() if is_torch_tensor(self.hashes) else self.hashes
        if self.width == 64:
            return np.ascontiguousarray(h).view(np.uint8).reshape(-1, 8).astype(np.uint32)
        return h.reshape(-1, 32).view("<u2").astype(np.uint32)

    def bucket(self, chunk: int, value: int) -> np.ndarray:
        """Dense ids whose `chunk` equals `value`, in input order (hamminghash.rs:133-138)."""
        if self._csr is None:
            ch = self._chunks()
            self._csr = []
            for k in range(self.num_chunks):
                order = np.argsort(ch[:, k], kind="stable").astype(np.uint32)
                offs = np.searchsorted(ch[order, k], np.arange(self.num_buckets + 1))
                self._csr.append((order, offs))
        order, offs = self._csr[chunk]
        return order[offs[value]:offs[value + 1]]


def find_groups(index: MIHIndex, max_dist: int, ctx=None):
    """hamminghash.rs:191-271 -> list of groups (seed first, then its unvisited neighbours).

    The device finds the exact adjacency; the reference's probing is complete for
    max_dist <= 31 (PDQ) / <= 15 (u64) and misses pairs above that.
    """
    ctx = ctx or default_context()
    n = len(index)
    if n == 0:
        return []
    members = np.empty(n, np.uint32)
    offsets = np.empty(n // 2 + 2, np.uint32)
    ng = C.c_size_t()
    h = index.hashes
    ctx.check(lib().rh_find_groups(ctx.handle, ptr(h), n, index.width, int(max_dist), ptr(members), members.size,
                                   ptr(offsets), offsets.size, C.byref(ng)))
    return [members[offsets[g]:offsets[g + 1]].tolist() for g in range(ng.value)]
