"""ctypes binding of librupphash_b200.so (the C ABI declared in include/rupphash_b200.h).

There is no CPU fallback: if the shared library is missing this module raises, and if no
CUDA device is usable `Context()` raises.  PyTorch is not needed by the library; CUDA
tensors are accepted only as a convenient source of device pointers.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("RH_B200_LIB") or os.path.join(_HERE, "librupphash_b200.so")   # RH_B200_LIB: A/B builds

RH_OK, RH_EINVAL, RH_ECUDA, RH_ENOMEM, RH_EUNSUPPORTED, RH_ENCCL = 0, -1, -2, -3, -4, -5
LAYOUT_RGB8, LAYOUT_RGBA8, LAYOUT_LUMA8 = 0, 1, 2
MAX_SIMILARITY_64 = 15   # hamminghash.rs:5
MAX_SIMILARITY_256 = 63  # hamminghash.rs:8
PDQ_MIN_QUALITY = 50     # scanner.rs:1579

# every symbol include/rupphash_b200.h declares (tests/test_abi.py checks the header against this)
EXPORTS = [
    "rh_ctx_create", "rh_ctx_destroy", "rh_ctx_set_stream", "rh_ctx_sync", "rh_ctx_set_option", "rh_last_error", "rh_version",
    "rh_kernel_launches", "rh_last_kernel_time", "rh_hamming_last_variant", "rh_alloc_pinned", "rh_free_pinned",
    "rh_pdq_hash_batch", "rh_pdq_hash_batch_async", "rh_ctx_wait", "rh_pdq_hash_from_coeffs", "rh_pdq_dihedral_from_coeffs", "rh_pdq_from_buffer64",
    "rh_phash_rotate_90", "rh_phash_rotate_180", "rh_phash_rotate_270", "rh_phash_flip_horizontal",
    "rh_phash_dihedral", "rh_phash_rotation_invariant", "rh_phash_batch",
    "rh_hamming_distances", "rh_hamming_distances_u64", "rh_hamming_group", "rh_hamming_group_shard",
    "rh_uf_merge", "rh_hamming_edges", "rh_hamming_group_u64", "rh_find_groups", "rh_group_max_dist",
    "rh_measure_peaks",
    "rh_group_create", "rh_group_destroy", "rh_group_size", "rh_group_ctx", "rh_group_last_error", "rh_group_info",
    "rh_group_last_times", "rh_hamming_group_multi", "rh_pdq_hash_batch_multi",
]


class RupphashError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"rupphash_b200 error {code}: {message}")
        self.code = code


class Unsupported(RupphashError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  rupphash_b200 has no CPU fallback.")
        L = C.CDLL(SO_PATH)
        _declare(L)
        _lib = L
    return _lib


_vp = C.c_void_p


def _declare(L):
    L.rh_ctx_create.argtypes = [C.c_int, C.POINTER(_vp)]
    L.rh_ctx_destroy.argtypes = [_vp]
    L.rh_ctx_set_stream.argtypes = [_vp, _vp]
    L.rh_ctx_sync.argtypes = [_vp]
    L.rh_ctx_set_option.argtypes = [_vp, C.c_char_p, C.c_int]
    L.rh_last_error.argtypes = [_vp]
    L.rh_last_error.restype = C.c_char_p
    L.rh_version.restype = C.c_char_p
    L.rh_kernel_launches.argtypes = [_vp]
    L.rh_kernel_launches.restype = C.c_uint64
    L.rh_last_kernel_time.argtypes = [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.rh_hamming_last_variant.argtypes = [_vp]
    L.rh_alloc_pinned.argtypes = [C.c_size_t, C.POINTER(_vp)]
    L.rh_free_pinned.argtypes = [_vp]
    L.rh_pdq_hash_batch.argtypes = [_vp, _vp, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                    _vp, _vp, _vp, _vp, _vp]
    L.rh_pdq_hash_batch_async.argtypes = [_vp, _vp, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                          _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_uint64)]
    L.rh_ctx_wait.argtypes = [_vp, C.c_uint64]
    L.rh_pdq_hash_from_coeffs.argtypes = [_vp, _vp, C.c_int64, _vp]
    L.rh_pdq_dihedral_from_coeffs.argtypes = [_vp, _vp, C.c_int64, _vp]
    L.rh_pdq_from_buffer64.argtypes = [_vp, _vp, C.c_int64, _vp, _vp, _vp, _vp]
    for name in ("rh_phash_rotate_90", "rh_phash_rotate_180", "rh_phash_rotate_270", "rh_phash_flip_horizontal",
                 "rh_phash_rotation_invariant"):
        getattr(L, name).argtypes = [C.c_uint64]
        getattr(L, name).restype = C.c_uint64
    L.rh_phash_dihedral.argtypes = [C.c_uint64, C.POINTER(C.c_uint64)]
    L.rh_phash_dihedral.restype = None
    L.rh_phash_batch.argtypes = [_vp, _vp, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_size_t, C.c_size_t, _vp, _vp]
    L.rh_hamming_distances.argtypes = [_vp, _vp, _vp, C.c_int64, _vp]
    L.rh_hamming_distances_u64.argtypes = [_vp, _vp, _vp, C.c_int64, _vp]
    L.rh_hamming_group.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_uint32, _vp, C.POINTER(C.c_uint64)]
    L.rh_hamming_group_shard.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_uint32, C.c_int, C.c_int, _vp,
                                         C.POINTER(C.c_uint64)]
    L.rh_uf_merge.argtypes = [_vp, _vp, C.c_int, C.c_int64, _vp]
    L.rh_hamming_edges.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_uint32, _vp, C.c_size_t,
                                   C.POINTER(C.c_uint64)]
    L.rh_hamming_group_u64.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_uint32, _vp,
                                       C.POINTER(C.c_uint64)]
    L.rh_find_groups.argtypes = [_vp, _vp, C.c_int64, C.c_int, C.c_uint32, _vp, C.c_size_t, _vp, C.c_size_t,
                                 C.POINTER(C.c_size_t)]
    L.rh_group_max_dist.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, _vp]
    L.rh_measure_peaks.argtypes = [_vp, C.POINTER(C.c_double)]
    L.rh_group_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_uint, C.POINTER(_vp)]
    L.rh_group_destroy.argtypes = [_vp]
    L.rh_group_size.argtypes = [_vp]
    L.rh_group_ctx.argtypes = [_vp, C.c_int]
    L.rh_group_ctx.restype = _vp
    L.rh_group_last_error.argtypes = [_vp]
    L.rh_group_last_error.restype = C.c_char_p
    L.rh_group_info.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.rh_group_last_times.argtypes = [_vp, C.POINTER(C.c_double), C.c_int]
    L.rh_hamming_group_multi.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_uint32, _vp,
                                         C.POINTER(C.c_uint64)]
    L.rh_pdq_hash_batch_multi.argtypes = [_vp, _vp, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                          _vp, _vp, _vp, _vp, _vp]


def is_torch_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def ptr(x):
    """Raw address of a numpy array (host) or torch tensor (host or CUDA); None stays None."""
    if x is None:
        return None
    if is_torch_tensor(x):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return x.ctypes.data
    raise TypeError(f"expected numpy array or torch tensor, got {type(x)}")


class Context:
    """One rh_ctx: bound to one CUDA device, one in-flight call at a time."""

    def __init__(self, device: int = 0):
        self._h = _vp()
        rc = lib().rh_ctx_create(int(device), C.byref(self._h))
        if rc != RH_OK:
            self._h = _vp()
            raise RupphashError(rc, f"rh_ctx_create(device={device}) failed: no usable CUDA device "
                                    "(rupphash_b200 has no CPU fallback)")
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().rh_ctx_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def check(self, rc: int):
        if rc == RH_OK:
            return
        msg = lib().rh_last_error(self._h).decode("utf-8", "replace")
        if rc == RH_EUNSUPPORTED:
            raise Unsupported(rc, msg)
        if rc == RH_EINVAL:
            raise ValueError(f"rupphash_b200: {msg}")
        raise RupphashError(rc, msg)

    def set_stream(self, cuda_stream_handle):
        self.check(lib().rh_ctx_set_stream(self._h, cuda_stream_handle))

    def sync(self):
        self.check(lib().rh_ctx_sync(self._h))

    def set_option(self, key: str, value: int):
        """Tuning knob for benchmarks / A-B runs (rh_ctx_set_option); defaults are the product path."""
        self.check(lib().rh_ctx_set_option(self._h, key.encode(), int(value)))

    @property
    def kernel_launches(self) -> int:
        return int(lib().rh_kernel_launches(self._h))

    def last_kernel_time(self):
        ms, units = C.c_double(), C.c_double()
        lib().rh_last_kernel_time(self._h, C.byref(ms), C.byref(units))
        return ms.value, units.value

    def hamming_last_variant(self) -> int:
        """tile-kernel variant of the last search on this ctx (0, 3..7; -1 before the first)"""
        return int(lib().rh_hamming_last_variant(self._h))

    def measure_peaks(self) -> dict:
        out = (C.c_double * 4)()
        self.check(lib().rh_measure_peaks(self._h, out))
        return {"popc_per_s": out[0], "lop3_per_s": out[1], "h2d_gbs": out[2], "copy_gbs": out[3]}


GROUP_NO_NCCL, GROUP_STATIC_TILES, GROUP_STEAL_TILES = 1, 2, 4


def _preload_bundled_nccl():
    """The library dlopens libnccl.so.2 at rh_group_create.  A Python process that imports torch AFTER that
    would find the (older) system copy already loaded under the same soname and fail to resolve torch's
    symbols, so the pip-bundled NCCL that torch uses is loaded first when it exists (RH_NCCL_LIB overrides)."""
    if os.environ.get("RH_NCCL_LIB") or "torch" in sys.modules:
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            path = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(path):
                C.CDLL(path, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass


class Group:
    """rh_group: several GPUs of one box driven from this one process (in-library NCCL / NVLink peer
    access; no torch.distributed involved)."""

    def __init__(self, devices=None, n_dev: int = 0, flags: int = 0):
        self._h = _vp()
        if not flags & GROUP_NO_NCCL:
            _preload_bundled_nccl()
        if devices is not None:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = lib().rh_group_create(arr, len(devices), int(flags), C.byref(self._h))
        else:
            rc = lib().rh_group_create(None, int(n_dev), int(flags), C.byref(self._h))
        if rc != RH_OK:
            self._h = _vp()
            raise RupphashError(rc, f"rh_group_create(devices={devices}, n_dev={n_dev}) failed "
                                    "(no usable CUDA devices, no peer access, or NCCL unavailable)")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().rh_group_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def size(self) -> int:
        return int(lib().rh_group_size(self._h))

    def info(self) -> dict:
        v, ws = C.c_int(), C.c_int()
        lib().rh_group_info(self._h, C.byref(v), C.byref(ws))
        return {"nccl_version": v.value, "work_stealing": bool(ws.value), "n_gpus": self.size}

    def last_times(self) -> dict:
        out = (C.c_double * 12)()
        lib().rh_group_last_times(self._h, out, 12)
        return {"group_wall_ms": out[0], "tile_ms_max": out[1], "tile_ms_min": out[2], "tile_ms_sum": out[3],
                "hash_wall_ms": out[4], "gpu0_inputs_ms": out[5], "gpu0_dense_arrays_ms": out[6],
                "gpu0_exchange_ms": out[7], "gpu0_merge_copyout_ms": out[8], "gpu0_timeline_ms": out[9]}

    def check(self, rc: int):
        if rc == RH_OK:
            return
        msg = lib().rh_group_last_error(self._h).decode("utf-8", "replace")
        if rc == RH_EINVAL:
            raise ValueError(f"rupphash_b200: {msg}")
        raise RupphashError(rc, msg)


_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    c = _default_ctx.get(device)
    if c is None:
        c = _default_ctx[device] = Context(device)
    return c


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """numpy array backed by page-locked memory (rh_alloc_pinned); freed with the array."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = _vp()
    rc = lib().rh_alloc_pinned(max(nbytes, 1), C.byref(p))
    if rc != RH_OK:
        raise MemoryError("rh_alloc_pinned failed")
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _pinned_keepalive[arr.ctypes.data] = p.value
    return arr


_pinned_keepalive: dict[int, int] = {}


def pinned_free(arr: np.ndarray):
    p = _pinned_keepalive.pop(arr.ctypes.data, None)
    if p is not None:
        lib().rh_free_pinned(p)
