"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(rupphash_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

LAYOUT_RGB8, LAYOUT_RGBA8, LAYOUT_LUMA8 = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_pdq.c", "oracle_group.c", "oracle_phash.c", "oracle.h", "Makefile")]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def _declare(L):
    L.orc_target_dimensions.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, _u32p, _u32p]
    L.orc_target_dimensions.restype = None
    L.orc_luma601.argtypes = [_u8p, C.c_int, C.c_size_t, _u8p]
    L.orc_luma601.restype = None
    L.orc_resize_box_u8.argtypes = [_u8p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int]
    L.orc_box_one_d.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]
    L.orc_box_one_d.restype = None
    L.orc_jarosz.argtypes = [_f32p] + [C.c_size_t] * 5
    L.orc_jarosz.restype = None
    L.orc_decimate64.argtypes = [_f32p, C.c_size_t, C.c_size_t, _f32p]
    L.orc_decimate64.restype = None
    L.orc_quality.argtypes = [_f32p, C.c_size_t, C.c_size_t]
    L.orc_quality.restype = C.c_float
    L.orc_dct_matrix.argtypes = [_f32p]
    L.orc_dct_matrix.restype = None
    L.orc_dct64_to_16.argtypes = [_f32p, _f32p]
    L.orc_dct64_to_16.restype = None
    L.orc_to_hash.argtypes = [_f32p, _u8p]
    L.orc_to_hash.restype = None
    L.orc_dihedral.argtypes = [_f32p, _u8p]
    L.orc_dihedral.restype = None
    L.orc_pdq_from_luma.argtypes = [_u8p, C.c_uint32, C.c_uint32, _f32p, _f32p, _f32p]
    L.orc_pdq_from_luma.restype = None
    L.orc_pdq_features.argtypes = [_u8p, C.c_int, C.c_uint32, C.c_uint32, _f32p, _f32p, _f32p]
    L.orc_quality_100.argtypes = [C.c_float]
    L.orc_quality_100.restype = C.c_uint16
    L.orc_pdq_batch_mt.argtypes = [_u8p, C.c_int, C.c_size_t, C.c_uint32, C.c_uint32, C.c_size_t, C.c_int,
                                   _u8p, _f32p, _f32p, _u8p, _u8p]
    for name in ("orc_phash_rot90", "orc_phash_rot180", "orc_phash_rot270", "orc_phash_flip_h", "orc_phash_rot_invariant"):
        getattr(L, name).argtypes = [C.c_uint64]
        getattr(L, name).restype = C.c_uint64
    L.orc_phash_dihedral.argtypes = [C.c_uint64, _u64p]
    L.orc_phash_dihedral.restype = None
    L.orc_phash_from_luma32.argtypes = [_u8p]
    L.orc_phash_from_luma32.restype = C.c_uint64
    L.orc_phash_image.argtypes = [_u8p, C.c_int, C.c_uint32, C.c_uint32, _u8p]
    L.orc_phash_image.restype = C.c_uint64
    L.orc_hamming256.argtypes = [_u8p, _u8p]
    L.orc_hamming256.restype = C.c_uint32
    L.orc_hamming64.argtypes = [C.c_uint64, C.c_uint64]
    L.orc_hamming64.restype = C.c_uint32
    L.orc_mih_new.argtypes = [_u8p, C.c_size_t, C.c_int]
    L.orc_mih_new.restype = C.c_void_p
    L.orc_mih_free.argtypes = [C.c_void_p]
    L.orc_mih_free.restype = None
    L.orc_mih_bucket.argtypes = [C.c_void_p, C.c_int, C.c_uint16, C.POINTER(C.c_size_t)]
    L.orc_mih_bucket.restype = _u32p
    L.orc_mih_offsets.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    L.orc_mih_offsets.restype = _u32p
    L.orc_find_groups.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.POINTER(_u32p), C.POINTER(_u32p), C.POINTER(C.c_size_t)]
    L.orc_free.argtypes = [C.c_void_p]
    L.orc_free.restype = None
    L.orc_group_generic.argtypes = [_u8p, _u8p, _u8p, _u8p, _u8p, C.c_size_t, C.c_uint32, C.c_int, C.c_int,
                                    _u32p, _u64p, _u32p, C.c_size_t]
    L.orc_group_tiles_rank.argtypes = [_u8p, _u8p, _u8p, _u8p, _u8p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int,
                                       C.c_int, _u32p, _u64p]
    L.orc_merge_parents.argtypes = [_u32p, C.c_int, C.c_size_t, _u32p]
    L.orc_merge_parents.restype = None


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


# ------------------------------------------------------------------ PDQ ----

def target_dimensions(w, h, max_dim=512):
    a, b = C.c_uint32(), C.c_uint32()
    lib().orc_target_dimensions(w, h, max_dim, C.byref(a), C.byref(b))
    return a.value, b.value


def luma601(px: np.ndarray, layout=LAYOUT_RGB8) -> np.ndarray:
    px = _c(px, np.uint8)
    ch = {LAYOUT_RGB8: 3, LAYOUT_RGBA8: 4, LAYOUT_LUMA8: 1}[layout]
    n = px.size // ch
    out = np.empty(n, np.uint8)
    lib().orc_luma601(_p(px, _u8p), layout, n, _p(out, _u8p))
    return out


def resize_box_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = _c(src, np.uint8)
    sh, sw = src.shape
    dst = np.empty((dh, dw), np.uint8)
    rc = lib().orc_resize_box_u8(_p(src, _u8p), sw, sh, _p(dst, _u8p), dw, dh)
    if rc:
        raise RuntimeError("orc_resize_box_u8 failed")
    return dst


def box_one_d(vec: np.ndarray, win: int) -> np.ndarray:
    vec = _c(vec, np.float32)
    out = np.zeros_like(vec)
    lib().orc_box_one_d(_p(vec, _f32p), 0, _p(out, _f32p), 0, vec.size, 1, win)
    return out


def jarosz(plane: np.ndarray, w_rows: int, w_cols: int, nreps: int = 2) -> np.ndarray:
    buf = np.array(plane, dtype=np.float32, order="C", copy=True)
    rows, cols = buf.shape
    lib().orc_jarosz(_p(buf, _f32p), rows, cols, w_rows, w_cols, nreps)
    return buf


def decimate64(plane: np.ndarray) -> np.ndarray:
    plane = _c(plane, np.float32)
    out = np.empty((64, 64), np.float32)
    lib().orc_decimate64(_p(plane, _f32p), plane.shape[0], plane.shape[1], _p(out, _f32p))
    return out


def quality(buf: np.ndarray) -> float:
    buf = _c(buf, np.float32)
    return float(lib().orc_quality(_p(buf, _f32p), buf.shape[0], buf.shape[1]))


def dct_matrix() -> np.ndarray:
    d = np.empty((16, 64), np.float32)
    lib().orc_dct_matrix(_p(d, _f32p))
    return d


def dct64_to_16(buf64: np.ndarray) -> np.ndarray:
    buf64 = _c(buf64, np.float32)
    out = np.empty(256, np.float32)
    lib().orc_dct64_to_16(_p(buf64, _f32p), _p(out, _f32p))
    return out


def to_hash(coeffs: np.ndarray) -> np.ndarray:
    coeffs = _c(coeffs, np.float32)
    out = np.empty(32, np.uint8)
    lib().orc_to_hash(_p(coeffs, _f32p), _p(out, _u8p))
    return out


def dihedral(coeffs: np.ndarray) -> np.ndarray:
    coeffs = _c(coeffs, np.float32)
    out = np.empty((8, 32), np.uint8)
    lib().orc_dihedral(_p(coeffs, _f32p), _p(out, _u8p))
    return out


def pdq_from_luma(luma: np.ndarray):
    """-> (coeffs[256], quality, buf64[64,64])"""
    luma = _c(luma, np.uint8)
    h, w = luma.shape
    coeffs = np.empty(256, np.float32)
    buf = np.empty((64, 64), np.float32)
    q = C.c_float()
    lib().orc_pdq_from_luma(_p(luma, _u8p), w, h, _p(coeffs, _f32p), C.byref(q), _p(buf, _f32p))
    return coeffs, q.value, buf


def pdq_features(img: np.ndarray, layout=LAYOUT_RGB8):
    """img: (h, w, ch) or (h, w) uint8 -> None | (coeffs, quality, buf64)"""
    img = _c(img, np.uint8)
    h, w = img.shape[:2]
    coeffs = np.empty(256, np.float32)
    buf = np.empty((64, 64), np.float32)
    q = C.c_float()
    rc = lib().orc_pdq_features(_p(img, _u8p), layout, w, h, _p(coeffs, _f32p), C.byref(q), _p(buf, _f32p))
    if rc:
        return None
    return coeffs, q.value, buf


def quality_100(q: float) -> int:
    return int(lib().orc_quality_100(q))


def pdq_batch(imgs: np.ndarray, layout=LAYOUT_RGB8, threads=1, want_coeffs=False, want_dihedral=False):
    """imgs: (n, h, w, ch) uint8 -> dict(hash (n,32), quality (n,), valid (n,), [coeffs], [dihedral])"""
    imgs = _c(imgs, np.uint8)
    n, h, w = imgs.shape[:3]
    pitch = imgs[0].nbytes if n else 0
    out = {
        "hash": np.zeros((n, 32), np.uint8),
        "quality": np.zeros(n, np.float32),
        "valid": np.zeros(n, np.uint8),
        "coeffs": np.zeros((n, 256), np.float32) if want_coeffs else None,
        "dihedral": np.zeros((n, 8, 32), np.uint8) if want_dihedral else None,
    }
    lib().orc_pdq_batch_mt(_p(imgs, _u8p), layout, n, w, h, pitch, threads, _p(out["hash"], _u8p),
                           _p(out["quality"], _f32p), _p(out["coeffs"], _f32p), _p(out["dihedral"], _u8p),
                           _p(out["valid"], _u8p))
    return out


# ---------------------------------------------------------------- pHash ----

def phash_rot90(h): return int(lib().orc_phash_rot90(h))
def phash_rot180(h): return int(lib().orc_phash_rot180(h))
def phash_rot270(h): return int(lib().orc_phash_rot270(h))
def phash_flip_h(h): return int(lib().orc_phash_flip_h(h))
def phash_rot_invariant(h): return int(lib().orc_phash_rot_invariant(h))


def phash_dihedral(h):
    out = (C.c_uint64 * 8)()
    lib().orc_phash_dihedral(h, out)
    return [int(x) for x in out]


def phash_from_luma32(luma: np.ndarray) -> int:
    luma = _c(luma, np.uint8)
    assert luma.size == 1024
    return int(lib().orc_phash_from_luma32(_p(luma, _u8p)))


def phash_image(img: np.ndarray, layout=LAYOUT_RGB8):
    img = _c(img, np.uint8)
    h, w = img.shape[:2]
    luma = np.empty((32, 32), np.uint8)
    v = lib().orc_phash_image(_p(img, _u8p), layout, w, h, _p(luma, _u8p))
    return int(v), luma


# -------------------------------------------------------------- Hamming ----

def hamming256(a, b) -> int:
    a, b = _c(a, np.uint8), _c(b, np.uint8)
    return int(lib().orc_hamming256(_p(a, _u8p), _p(b, _u8p)))


def hamming64(a: int, b: int) -> int:
    return int(lib().orc_hamming64(a, b))


class MIHIndex:
    """hamminghash.rs:82-149"""

    def __init__(self, hashes: np.ndarray):
        if hashes.dtype == np.uint64:
            self.width = 64
            raw = np.ascontiguousarray(hashes).view(np.uint8)
            self.n = hashes.size
        else:
            self.width = 256
            raw = _c(hashes, np.uint8).reshape(-1, 32)
            self.n = raw.shape[0]
        self._raw = raw
        self._h = lib().orc_mih_new(_p(raw, _u8p), self.n, self.width)
        if not self._h:
            raise RuntimeError("orc_mih_new failed")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_mih_free(self._h)
            self._h = None

    def __len__(self):
        return self.n

    def bucket(self, chunk: int, value: int) -> np.ndarray:
        ln = C.c_size_t()
        p = lib().orc_mih_bucket(self._h, chunk, value, C.byref(ln))
        return np.ctypeslib.as_array(p, shape=(ln.value,)).copy() if ln.value else np.empty(0, np.uint32)

    def find_groups(self, max_dist: int, threads: int = 1):
        """hamminghash.rs:191-271 -> list of lists (seed first)"""
        mem, off = _u32p(), _u32p()
        ng = C.c_size_t()
        lib().orc_find_groups(self._h, max_dist, threads, C.byref(mem), C.byref(off), C.byref(ng))
        offs = np.ctypeslib.as_array(off, shape=(ng.value + 1,)).copy()
        total = int(offs[-1])
        mems = np.ctypeslib.as_array(mem, shape=(max(total, 1),)).copy()
        lib().orc_free(mem)
        lib().orc_free(off)
        return [mems[offs[g]:offs[g + 1]].tolist() for g in range(ng.value)]


def group_generic(hashes, similarity, has_hash=None, variants=None, n_variants=None, low_conf=None,
                  threads=1, use_mih=True, edges_cap=0):
    """scanner.rs:1640-1817 -> (labels[n] = min index of component, edge_count, edges[k,2] or None)"""
    hashes = _c(hashes, np.uint8).reshape(-1, 32)
    n = hashes.shape[0]
    has_hash, variants = _c(has_hash, np.uint8), _c(variants, np.uint8)
    n_variants, low_conf = _c(n_variants, np.uint8), _c(low_conf, np.uint8)
    labels = np.empty(n, np.uint32)
    cnt = C.c_uint64()
    edges = np.zeros((edges_cap, 2), np.uint32) if edges_cap else None
    rc = lib().orc_group_generic(_p(hashes, _u8p), _p(has_hash, _u8p), _p(variants, _u8p), _p(n_variants, _u8p),
                                 _p(low_conf, _u8p), n, similarity, threads, 1 if use_mih else 0,
                                 _p(labels, _u32p), C.byref(cnt), _p(edges, _u32p), edges_cap)
    if rc:
        raise ValueError("similarity above 63 is not supported (scanner.rs:1650-1655)")
    if edges is not None:
        edges = edges[: min(edges_cap, cnt.value)]
    return labels, int(cnt.value), edges


def group_generic_sampled(hashes, similarity, chunk_stride, has_hash=None, variants=None, n_variants=None,
                          low_conf=None, threads=1, sample_files=0):
    """bench.py timing aid (inputs whose full CPU search takes minutes): index build as usual, probes only for
    every chunk_stride-th 2000-file chunk of query files -> edge count of that sample."""
    hashes = _c(hashes, np.uint8).reshape(-1, 32)
    n = hashes.shape[0]
    has_hash, variants = _c(has_hash, np.uint8), _c(variants, np.uint8)
    n_variants, low_conf = _c(n_variants, np.uint8), _c(low_conf, np.uint8)
    labels = np.empty(n, np.uint32)
    cnt = C.c_uint64()
    L = lib()
    L.orc_group_generic_sampled.argtypes = [_u8p, _u8p, _u8p, _u8p, _u8p, C.c_size_t, C.c_uint32, C.c_int, C.c_size_t,
                                            C.c_size_t, _u32p, _u64p]
    rc = L.orc_group_generic_sampled(_p(hashes, _u8p), _p(has_hash, _u8p), _p(variants, _u8p), _p(n_variants, _u8p),
                                     _p(low_conf, _u8p), n, similarity, threads, int(chunk_stride), int(sample_files),
                                     _p(labels, _u32p), C.byref(cnt))
    if rc:
        raise ValueError("similarity above 63 is not supported (scanner.rs:1650-1655)")
    return int(cnt.value)


def group_tiles_rank(hashes, similarity, tile, rank, world, has_hash=None, variants=None, n_variants=None, low_conf=None):
    hashes = _c(hashes, np.uint8).reshape(-1, 32)
    n = hashes.shape[0]
    has_hash, variants = _c(has_hash, np.uint8), _c(variants, np.uint8)
    n_variants, low_conf = _c(n_variants, np.uint8), _c(low_conf, np.uint8)
    parent = np.empty(n, np.uint32)
    cnt = C.c_uint64()
    rc = lib().orc_group_tiles_rank(_p(hashes, _u8p), _p(has_hash, _u8p), _p(variants, _u8p), _p(n_variants, _u8p),
                                    _p(low_conf, _u8p), n, similarity, tile, rank, world, _p(parent, _u32p), C.byref(cnt))
    if rc:
        raise ValueError("bad arguments")
    return parent, int(cnt.value)


def merge_parents(parents: np.ndarray) -> np.ndarray:
    parents = _c(parents, np.uint32)
    world, n = parents.shape
    out = np.empty(n, np.uint32)
    lib().orc_merge_parents(_p(parents, _u32p), world, n, _p(out, _u32p))
    return out


def labels_to_groups(labels: np.ndarray):
    """canonical groups: members ascending, groups ordered by first member, len > 1 (scanner.rs:1817)"""
    order = np.argsort(labels, kind="stable")
    sl = labels[order]
    groups = []
    start = 0
    for i in range(1, len(sl) + 1):
        if i == len(sl) or sl[i] != sl[start]:
            if i - start > 1:
                groups.append(order[start:i].tolist())
            start = i
    return groups
