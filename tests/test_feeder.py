"""scanner.hash_files_batched, host logic with an injected hasher: batching by shape,
arrival order, None for unhashable images, progress ticks, error propagation, decode/hash overlap."""
import threading
import time

import numpy as np
import pytest

from rupphash_b200 import scanner


def fake_hasher(calls, delay=0.0):
    def h(arr):
        calls.append((arr.shape, threading.current_thread().name))
        if delay:
            time.sleep(delay)
        n = arr.shape[0]
        first = arr.reshape(n, -1)[:, 0]                       # the first pixel identifies the image
        return {"hash": np.repeat(first[:, None], 32, axis=1).astype(np.uint8),
                "quality": (first.astype(np.float32) / 255.0), "valid": (first != 13).astype(np.uint8),
                "coeffs": np.repeat(first[:, None].astype(np.float32), 256, axis=1)}
    return h


def images(shapes):
    for k, s in enumerate(shapes):
        yield np.full(s, k, np.uint8)


def test_batches_by_shape_and_keeps_arrival_order():
    shapes = [(8, 8, 3), (6, 9, 3), (8, 8, 3), (4, 100, 3), (8, 8, 3), (6, 9, 3), (8, 8, 3)]
    calls, ticks = [], []
    res = scanner.hash_files_batched(images(shapes), batch_size=2, hasher=fake_hasher(calls),
                                     progress=lambda d, t: ticks.append((d, t)))
    assert len(res) == len(shapes)
    assert res[3] is None                                      # narrower than 5 px: pdqhash.rs:167-169
    for k in (0, 1, 2, 4, 5, 6):
        assert res[k]["hash"][0] == k and res[k]["coeffs"][0] == k
        assert res[k]["quality_100"] == scanner.quality_100(k / 255.0)
    # full batches of two, then the leftovers; every call on the one submitter thread
    assert sorted(c[0] for c in calls) == sorted([(2, 8, 8, 3), (2, 8, 8, 3), (2, 6, 9, 3)])
    assert {c[1] for c in calls} == {"rh-submitter"}
    assert [d for d, _ in ticks] == sorted(d for d, _ in ticks) and ticks[-1][0] == len(shapes)


def test_invalid_images_of_a_batch_become_none():
    shapes = [(8, 8, 3)] * 20
    res = scanner.hash_files_batched(images(shapes), batch_size=8, hasher=fake_hasher([]))
    assert res[13] is None and all(r is not None for k, r in enumerate(res) if k != 13)


def test_empty_input():
    assert scanner.hash_files_batched(iter([]), hasher=fake_hasher([])) == []


def test_hasher_error_reaches_the_caller():
    def boom(arr):
        raise RuntimeError("device lost")
    with pytest.raises(RuntimeError, match="device lost"):
        scanner.hash_files_batched(images([(8, 8, 3)] * 50), batch_size=4, hasher=boom)


def test_decode_and_hashing_overlap():
    """A slow iterator (decode) and a slow hasher run side by side: the wall time is far below the sum."""
    n, bs, t_decode, t_hash = 24, 4, 0.02, 0.08

    def slow_images():
        for k in range(n):
            time.sleep(t_decode)
            yield np.full((8, 8, 3), k, np.uint8)
    t0 = time.perf_counter()
    res = scanner.hash_files_batched(slow_images(), batch_size=bs, hasher=fake_hasher([], delay=t_hash))
    wall = time.perf_counter() - t0
    serial = n * t_decode + (n // bs) * t_hash
    assert len(res) == n and wall < 0.8 * serial, (wall, serial)


def test_batches_are_bounded_by_bytes():
    """a batch never stages more than batch_bytes of pixels, whatever batch_size says."""
    calls = []
    shapes = [(100, 100, 3)] * 7
    res = scanner.hash_files_batched(images(shapes), batch_size=256, batch_bytes=70_000, hasher=fake_hasher(calls))
    assert [r["hash"][0] for r in res] == list(range(7))
    assert sorted(c[0][0] for c in calls) == [1, 2, 2, 2]          # 30 kB images: two per batch


def test_open_shape_slots_are_capped():
    """Many camera resolutions: at most max_open_shapes partly filled batches exist, the least recently used
    one is hashed early; nothing is lost or reordered."""
    shapes = [(8 + (k % 5), 8, 3) for k in range(40)]
    calls = []
    res = scanner.hash_files_batched(images(shapes), batch_size=64, max_open_shapes=2, hasher=fake_hasher(calls))
    assert res[13] is None                                          # the fake hasher's "invalid" image
    assert [r["hash"][0] for k, r in enumerate(res) if k != 13] == [k for k in range(40) if k != 13]
    assert len(calls) > 5                                           # evictions produced early, partial batches
    assert sum(c[0][0] for c in calls) == 40


def test_decode_workers_fill_batches_in_parallel():
    """A pool of decode threads (the reference's rayon pool, scanner.rs:1188-1205) pulls file names, decodes and
    stages; results still come back in arrival order and decoding overlaps itself."""
    n, t_decode = 48, 0.02
    names = []

    def decode(k):
        names.append(threading.current_thread().name)
        time.sleep(t_decode)
        return None if k == 7 else np.full((8, 8, 3), k, np.uint8)
    t0 = time.perf_counter()
    res = scanner.hash_files_batched(range(n), decode=decode, workers=8, batch_size=4, hasher=fake_hasher([]))
    wall = time.perf_counter() - t0
    assert res[7] is None and res[13] is None
    assert [r["hash"][0] for k, r in enumerate(res) if r is not None] == [k for k in range(n) if k not in (7, 13)]
    assert len(set(names)) > 1 and all(nm.startswith("rh-decode-") for nm in names)
    assert wall < 0.5 * n * t_decode, wall


def test_decode_error_reaches_the_caller():
    def decode(k):
        if k == 5:
            raise OSError("unreadable file")
        return np.full((8, 8, 3), k, np.uint8)
    with pytest.raises(OSError, match="unreadable"):
        scanner.hash_files_batched(range(20), decode=decode, workers=3, batch_size=4, hasher=fake_hasher([]))
