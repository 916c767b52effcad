"""Host mirror of src/phash.rs: the bit-level dihedral operations (phash.rs:137-255, exact) and
DctPhash (phash.rs:33-89) over the CUDA library."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import default_context, is_torch_tensor, lib, ptr
from .pdqhash import _empty_like_source, _layout_of


def rotate_hash_90(h: int) -> int:
    return int(lib().rh_phash_rotate_90(h))


def rotate_hash_180(h: int) -> int:
    return int(lib().rh_phash_rotate_180(h))


def rotate_hash_270(h: int) -> int:
    return int(lib().rh_phash_rotate_270(h))


def flip_hash_horizontal(h: int) -> int:
    return int(lib().rh_phash_flip_horizontal(h))


def generate_dihedral_hashes(h: int) -> list:
    out = (C.c_uint64 * 8)()
    lib().rh_phash_dihedral(h, out)
    return [int(x) for x in out]


def calculate_rotation_invariant_hash(h: int) -> int:
    return int(lib().rh_phash_rotation_invariant(h))


class DctPhash:
    """phash.rs:20-89"""

    def __init__(self, ctx=None):
        self._ctx = ctx

    @classmethod
    def new(cls):
        return cls()

    def hash_batch(self, images, want_dihedral=False):
        ctx = self._ctx or default_context()
        if not is_torch_tensor(images):
            images = np.ascontiguousarray(images, dtype=np.uint8)
        n, h, w = images.shape[:3]
        layout = _layout_of(tuple(images.shape[1:]))
        out = _empty_like_source(images, (n,), np.uint64)
        dih = _empty_like_source(images, (n, 8), np.uint64) if want_dihedral else None
        ctx.check(lib().rh_phash_batch(ctx.handle, ptr(images), layout, n, w, h, 0, 0, ptr(out), ptr(dih)))
        return (out, dih) if want_dihedral else out

    def hash_image(self, image) -> int:
        return int(self.hash_batch(np.ascontiguousarray(image, dtype=np.uint8)[None])[0])

    def hash_image_invariant(self, image) -> int:
        return calculate_rotation_invariant_hash(self.hash_image(image))
