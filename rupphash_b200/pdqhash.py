"""Host mirror of src/pdqhash.rs's public API over the CUDA library.

    generate_pdq_features(image)   pdqhash.rs:166-196   -> None | (PdqFeatures, quality)
    generate_pdq(image)            pdqhash.rs:199-201   -> None | (hash[32], quality)
    PdqFeatures.to_hash            pdqhash.rs:59-61
    PdqFeatures.generate_dihedral_hashes  pdqhash.rs:71-87
    hash_batch(images)             the batched form the scanner feeds (scanner.rs:1409-1418)

Images are numpy arrays (h, w, 3) RGB8 / (h, w, 4) RGBA8 / (h, w) Luma8 -- what
image::DynamicImage holds after decode -- or CUDA tensors of the same shapes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import LAYOUT_LUMA8, LAYOUT_RGB8, LAYOUT_RGBA8, default_context, is_torch_tensor, lib, ptr

MIN_HASHABLE_DIM = 5     # pdqhash.rs:17
DOWNSAMPLE_DIMS = 512    # pdqhash.rs:19
HASH_LENGTH = 32         # pdqhash.rs:23


def _layout_of(shape) -> int:
    if len(shape) == 2:
        return LAYOUT_LUMA8
    if len(shape) == 3 and shape[2] == 3:
        return LAYOUT_RGB8
    if len(shape) == 3 and shape[2] == 4:
        return LAYOUT_RGBA8
    if len(shape) == 3 and shape[2] == 1:
        return LAYOUT_LUMA8
    raise ValueError(f"unsupported image shape {tuple(shape)}")


def _empty_like_source(src, shape, dtype):
    """Output buffer on the same side as the input (CUDA tensor in -> CUDA tensors out)."""
    if is_torch_tensor(src) and src.is_cuda:
        import torch
        tdt = {np.uint8: torch.uint8, np.float32: torch.float32, np.uint32: torch.int32, np.uint64: torch.int64}[dtype]
        return torch.empty(shape, dtype=tdt, device=src.device)
    return np.empty(shape, dtype=dtype)


class PdqFeatures:
    """pdqhash.rs:48-51 -- the 16 x 16 DCT block, row-major (coefficients[16*r + c])."""

    __slots__ = ("coefficients", "_ctx")

    def __init__(self, coefficients, ctx=None):
        self.coefficients = np.ascontiguousarray(coefficients, dtype=np.float32).reshape(256)
        self._ctx = ctx

    def to_hash(self) -> np.ndarray:
        ctx = self._ctx or default_context()
        out = np.empty(32, np.uint8)
        ctx.check(lib().rh_pdq_hash_from_coeffs(ctx.handle, ptr(self.coefficients), 1, ptr(out)))
        return out

    def generate_dihedral_hashes(self) -> np.ndarray:
        """(8, 32): identity, rot90, rot180, rot270, mirror-x, mirror-y, transpose, anti-transpose."""
        ctx = self._ctx or default_context()
        out = np.empty((8, 32), np.uint8)
        ctx.check(lib().rh_pdq_dihedral_from_coeffs(ctx.handle, ptr(self.coefficients), 1, ptr(out)))
        return out


def dihedral_from_coeffs(coeffs, ctx=None):
    """Batched generate_dihedral_hashes: (n, 256) f32 -> (n, 8, 32) u8 (scanner.rs:1622)."""
    ctx = ctx or default_context()
    n = coeffs.shape[0]
    out = _empty_like_source(coeffs, (n, 8, 32), np.uint8)
    ctx.check(lib().rh_pdq_dihedral_from_coeffs(ctx.handle, ptr(coeffs), n, ptr(out)))
    return out


def hash_from_coeffs(coeffs, ctx=None):
    ctx = ctx or default_context()
    n = coeffs.shape[0]
    out = _empty_like_source(coeffs, (n, 32), np.uint8)
    ctx.check(lib().rh_pdq_hash_from_coeffs(ctx.handle, ptr(coeffs), n, ptr(out)))
    return out


def hash_batch(images, want_coeffs=False, want_dihedral=False, ctx=None, layout=None):
    """Batched generate_pdq_features + to_hash over n same-sized images.

    images: (n, h, w[, ch]) uint8, numpy (host) or CUDA tensor (device resident).
    Returns dict(hash (n,32), quality (n,), valid (n,), [coeffs (n,256)], [dihedral (n,8,32)])
    on the same side as the input.
    """
    ctx = ctx or default_context()
    if not is_torch_tensor(images):
        images = np.ascontiguousarray(images, dtype=np.uint8)
    shape = tuple(images.shape)
    if len(shape) < 3:
        raise ValueError("hash_batch expects (n, h, w[, ch])")
    n, h, w = shape[0], shape[1], shape[2]
    if layout is None:
        layout = _layout_of(shape[1:])
    out = {
        "hash": _empty_like_source(images, (n, 32), np.uint8),
        "quality": _empty_like_source(images, (n,), np.float32),
        "valid": _empty_like_source(images, (n,), np.uint8),
        "coeffs": _empty_like_source(images, (n, 256), np.float32) if want_coeffs else None,
        "dihedral": _empty_like_source(images, (n, 8, 32), np.uint8) if want_dihedral else None,
    }
    ctx.check(lib().rh_pdq_hash_batch(ctx.handle, ptr(images), layout, n, w, h, 0, 0, ptr(out["hash"]),
                                      ptr(out["quality"]), ptr(out["coeffs"]), ptr(out["dihedral"]),
                                      ptr(out["valid"])))
    return out


def generate_pdq_features(image, ctx=None):
    """pdqhash.rs:166-196: None when width or height < 5, else (PdqFeatures, quality in [0,1])."""
    image = np.ascontiguousarray(image, dtype=np.uint8)
    h, w = image.shape[:2]
    if w < MIN_HASHABLE_DIM or h < MIN_HASHABLE_DIM:
        return None
    out = hash_batch(image[None], want_coeffs=True, ctx=ctx)
    return PdqFeatures(out["coeffs"][0], ctx), float(out["quality"][0])


def generate_pdq(image, ctx=None):
    """pdqhash.rs:199-201"""
    image = np.ascontiguousarray(image, dtype=np.uint8)
    h, w = image.shape[:2]
    if w < MIN_HASHABLE_DIM or h < MIN_HASHABLE_DIM:
        return None
    out = hash_batch(image[None], ctx=ctx)
    return out["hash"][0].copy(), float(out["quality"][0])


def from_buffer64(buf64, want_coeffs=True, want_dihedral=False, ctx=None):
    """The tail alone (pdqhash.rs:258-260) over (n, 64, 64) f32 buffers."""
    ctx = ctx or default_context()
    if not is_torch_tensor(buf64):
        buf64 = np.ascontiguousarray(buf64, dtype=np.float32)
    n = buf64.shape[0]
    out = {
        "hash": _empty_like_source(buf64, (n, 32), np.uint8),
        "quality": _empty_like_source(buf64, (n,), np.float32),
        "coeffs": _empty_like_source(buf64, (n, 256), np.float32) if want_coeffs else None,
        "dihedral": _empty_like_source(buf64, (n, 8, 32), np.uint8) if want_dihedral else None,
    }
    ctx.check(lib().rh_pdq_from_buffer64(ctx.handle, ptr(buf64), n, ptr(out["hash"]), ptr(out["quality"]),
                                         ptr(out["coeffs"]), ptr(out["dihedral"])))
    return out
