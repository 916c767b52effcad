"""rh_group: several GPUs driven from one process through the C ABI (in-library NCCL over NVLink, tiles
claimed from one pool).  Parity bar: labels and comparison_count bit-identical to the CPU oracle and to
the single-GPU search, for every exchange / scheduling mode.  On a 1-GPU box the group has one device
(the multi-device cases are skipped); `gpurun --gpus 2|4|8` exercises the rest."""
import os
import struct
import subprocess

import numpy as np
import pytest

from rupphash_b200.synth import planted_hashes, random_variants

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _case(n, seed, variants=True):
    hashes, low_conf = planted_hashes(n, seed=seed)
    var = random_variants(hashes, seed=seed + 1) if variants else None
    return hashes, low_conf, var


def test_group_of_one_matches_oracle(orc):
    from rupphash_b200 import _lib, scanner
    g = _lib.Group(devices=[0])
    try:
        hashes, low_conf, var = _case(20_000, 5)
        want, want_cnt, _ = orc.group_generic(hashes, 31, variants=var, low_conf=low_conf, threads=4)
        labels, cnt = scanner.group_labels_multi(g, hashes, 31, variants=var, low_conf=low_conf)
        assert cnt == want_cnt and np.array_equal(labels, want)
        assert g.info()["n_gpus"] == 1
    finally:
        g.close()


@pytest.mark.parametrize("flags", [0, 1, 4, 5])   # NCCL + static tiles, peer copies, work stealing, both
@pytest.mark.parametrize("similarity", [31, 40])
def test_group_all_gpus_matches_oracle(orc, flags, similarity):
    from rupphash_b200 import _lib, scanner
    ng = _n_gpus()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus N)")
    g = _lib.Group(n_dev=ng, flags=flags)
    try:
        info = g.info()
        assert info["n_gpus"] == ng
        assert (info["nccl_version"] > 0) == (not flags & _lib.GROUP_NO_NCCL)
        rng = np.random.default_rng(3)
        n = 60_000
        hashes, low_conf, var = _case(n, 11 + similarity)
        has_hash = (rng.random(n) > 0.1).astype(np.uint8)
        want, want_cnt, _ = orc.group_generic(hashes, similarity, has_hash=has_hash, variants=var, low_conf=low_conf,
                                              threads=8)
        for rep in range(2):   # the second call reuses every buffer
            labels, cnt = scanner.group_labels_multi(g, hashes, similarity, has_hash=has_hash, variants=var,
                                                     low_conf=low_conf)
            assert cnt == want_cnt, (flags, rep)
            assert np.array_equal(labels, want), (flags, rep)
        # plain hashes, no variants (configs[2] shape of input)
        want2, cnt2, _ = orc.group_generic(hashes, similarity, low_conf=low_conf, threads=8)
        labels, cnt = scanner.group_labels_multi(g, hashes, similarity, low_conf=low_conf)
        assert cnt == cnt2 and np.array_equal(labels, want2)
        t = g.last_times()
        assert t["tile_ms_max"] > 0 and t["group_wall_ms"] >= t["tile_ms_max"]
    finally:
        g.close()


def test_group_device_resident_inputs(orc):
    """Inputs that already live on one GPU of the group (e.g. written there by the hashing kernels) are
    broadcast from it; labels may be written to the first GPU's memory."""
    import torch
    from rupphash_b200 import _lib, scanner
    ng = _n_gpus()
    g = _lib.Group(n_dev=ng)
    try:
        hashes, low_conf, var = _case(30_000, 21)
        want, want_cnt, _ = orc.group_generic(hashes, 31, variants=var, low_conf=low_conf, threads=4)
        src = torch.device("cuda", ng - 1)
        d_h, d_l, d_v = (torch.from_numpy(x).to(src) for x in (hashes, low_conf, var))
        out = torch.empty(len(hashes), dtype=torch.int32, device="cuda:0")
        torch.cuda.synchronize()
        labels, cnt = scanner.group_labels_multi(g, d_h, 31, variants=d_v, low_conf=d_l, out=out)
        assert cnt == want_cnt
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want)
    finally:
        g.close()


def test_group_rejects_bad_arguments():
    from rupphash_b200 import _lib, scanner
    g = _lib.Group(devices=[0])
    try:
        h = np.zeros((10, 32), np.uint8)
        with pytest.raises(ValueError):
            scanner.group_labels_multi(g, h, 64)          # scanner.rs:1650-1655
        labels, cnt = scanner.group_labels_multi(g, h[:0], 31)
        assert cnt == 0 and len(labels) == 0
    finally:
        g.close()
    with pytest.raises(_lib.RupphashError):
        _lib.Group(devices=[0, 0])


def test_hash_batch_multi_matches_single(orc):
    import ctypes as C
    from rupphash_b200 import _lib, pdqhash
    from rupphash_b200.synth import synth_images
    ng = _n_gpus()
    g = _lib.Group(n_dev=ng)
    try:
        imgs = synth_images(13, 768, 1024, seed=3)
        n = len(imgs)
        h = np.zeros((n, 32), np.uint8)
        q = np.zeros(n, np.float32)
        d = np.zeros((n, 8, 32), np.uint8)
        v = np.zeros(n, np.uint8)
        g.check(_lib.lib().rh_pdq_hash_batch_multi(g.handle, _lib.ptr(imgs), _lib.LAYOUT_RGB8, n, 1024, 768, 0, 0,
                                                   _lib.ptr(h), _lib.ptr(q), None, _lib.ptr(d), _lib.ptr(v)))
        want = orc.pdq_batch(imgs, threads=4, want_dihedral=True)
        assert np.array_equal(h, want["hash"]) and np.array_equal(q, want["quality"])
        assert np.array_equal(d, want["dihedral"]) and v.all()
    finally:
        g.close()


def _build_harness():
    import __graft_entry__ as ge
    from rupphash_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        ge.build()
    exe = os.path.join(ROOT, "tests", "cpp", "build", "group_multi_harness")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    so_dir = os.path.join(ROOT, "rupphash_b200")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", os.path.join(ROOT, "tests", "cpp", "group_multi_harness.cpp"),
                    "-o", exe, "-L" + so_dir, "-lrupphash_b200", "-Wl,-rpath," + so_dir], check=True,
                   capture_output=True)
    return exe


def run_harness(tmp_path, hashes, similarity, variants=None, low_conf=None, has_hash=None, n_gpus=0, flags=0):
    exe = _build_harness()
    fin, fout = os.path.join(tmp_path, "in.bin"), os.path.join(tmp_path, "out.bin")
    n = len(hashes)
    with open(fin, "wb") as f:
        f.write(struct.pack("<q8B", n, int(variants is not None), int(low_conf is not None), int(has_hash is not None),
                            0, 0, 0, 0, 0))
        f.write(np.ascontiguousarray(hashes, np.uint8).tobytes())
        for a in (variants, low_conf, has_hash):
            if a is not None:
                f.write(np.ascontiguousarray(a, np.uint8).tobytes())
    r = subprocess.run([exe, fin, fout, str(similarity), str(n_gpus), str(flags)], capture_output=True, text=True)
    assert r.returncode == 0, f"harness exit {r.returncode}: {r.stdout}{r.stderr}"
    raw = open(fout, "rb").read()
    edges, wall, tile_max, tile_sum, ng, nccl = struct.unpack_from("<QdddII", raw, 0)
    labels = np.frombuffer(raw, np.uint32, count=2 * n, offset=40)
    return {"edges": edges, "wall_ms": wall, "tile_ms_max": tile_max, "tile_ms_sum": tile_sum, "n_gpus": ng,
            "nccl": nccl, "multi": labels[:n], "single": labels[n:], "stdout": r.stdout}


def test_cpp_harness_groups_on_all_gpus(orc, tmp_path):
    """No Python and no torch inside the process that drives the GPUs: a C++ binary over the C ABI."""
    n = 200_000
    hashes, low_conf = planted_hashes(n, seed=0xB200, n_clusters=2000, identical_block=500)
    var = random_variants(hashes, seed=9)
    want, want_cnt, _ = orc.group_generic(hashes, 31, variants=var, low_conf=low_conf, threads=os.cpu_count() or 4)
    res = run_harness(str(tmp_path), hashes, 31, variants=var, low_conf=low_conf)
    assert res["edges"] == want_cnt
    assert np.array_equal(res["multi"], want) and np.array_equal(res["single"], want)
    assert res["n_gpus"] == _n_gpus()
