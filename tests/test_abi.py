"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/rupphash_b200.h declares, refuses to run without a device (no CPU fallback), and its
host-only pHash bit operations match the oracle."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rlib():
    import __graft_entry__ as g
    from rupphash_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        g.build()
    return _lib


def header_functions():
    text = open(os.path.join(ROOT, "include", "rupphash_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rh_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(rlib):
    names = header_functions()
    assert len(names) >= 25
    L = rlib.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(rlib.EXPORTS) == names, "rupphash_b200/_lib.py EXPORTS out of sync with the header"


def test_no_torch_in_library(rlib):
    import subprocess
    out = subprocess.run(["ldd", rlib.SO_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "c10" not in out


def test_no_cpu_fallback_without_device(rlib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(rlib.RupphashError):
        rlib.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rupphash_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle.h" not in text, f


def test_phash_bit_ops_match_oracle(rlib, orc):
    from rupphash_b200 import phash
    rng = np.random.default_rng(5)
    vals = [0, 1, 0xFFFFFFFFFFFFFFFF, 0x8000000000000000, 0x0123456789ABCDEF] + \
        [int(x) for x in rng.integers(0, 2**63, size=300, dtype=np.uint64)]
    for h in vals:
        assert phash.rotate_hash_90(h) == orc.phash_rot90(h)
        assert phash.rotate_hash_180(h) == orc.phash_rot180(h)
        assert phash.rotate_hash_270(h) == orc.phash_rot270(h)
        assert phash.flip_hash_horizontal(h) == orc.phash_flip_h(h)
        assert phash.generate_dihedral_hashes(h) == orc.phash_dihedral(h)
        assert phash.calculate_rotation_invariant_hash(h) == orc.phash_rot_invariant(h)


def test_quality_100_matches_oracle(orc):
    from rupphash_b200 import scanner
    for q in [0.0, 0.004, 0.005, 0.0049999, 0.495, 0.5, 0.505, 0.994999, 0.995, 1.0, 0.125, 0.745]:
        assert scanner.quality_100(q) == orc.quality_100(q), q
    rng = np.random.default_rng(1)
    for q in rng.random(500).astype(np.float32):
        assert scanner.quality_100(float(q)) == orc.quality_100(float(q))


def test_luma_division_magic():
    # floor(v / 1000) == (v * 4294968) >> 32 for every value the luma sum can take (pdq.cu luma601)
    v = np.arange(0, 255 * 1000 + 501, dtype=np.uint64)
    assert np.array_equal((v * np.uint64(4294968)) >> np.uint64(32), v // np.uint64(1000))


def test_labels_to_groups_canonical():
    from rupphash_b200 import scanner
    labels = np.array([0, 1, 0, 3, 1, 5, 3, 0], np.uint32)
    assert scanner.labels_to_groups(labels) == [[0, 2, 7], [1, 4], [3, 6]]


def test_mih_index_bucket_matches_oracle(orc):
    from rupphash_b200 import hamminghash
    rng = np.random.default_rng(3)
    h = rng.integers(0, 256, size=(500, 32), dtype=np.uint8)
    h[:, 0:2] = rng.integers(0, 3, size=(500, 2))  # force collisions in chunk 0
    mine, ref = hamminghash.MIHIndex.new(h), orc.MIHIndex(h)
    assert len(mine) == 500
    for chunk in (0, 7, 15):
        for value in {hamminghash.get_chunk(h[i], chunk) for i in range(0, 500, 37)} | {0, 65535}:
            assert mine.bucket(chunk, value).tolist() == ref.bucket(chunk, value).tolist()
    u = rng.integers(0, 2**63, size=300, dtype=np.uint64)
    mine, ref = hamminghash.MIHIndex.new(u), orc.MIHIndex(u)
    for chunk in range(8):
        v = hamminghash.get_chunk(int(u[5]), chunk)
        assert mine.bucket(chunk, v).tolist() == ref.bucket(chunk, v).tolist()


def test_merge_groups_by_stem():
    """scanner.rs:1905-1983: groups sharing a (directory, stem) pair are merged, output canonical."""
    from rupphash_b200 import scanner
    paths = ["/a/IMG_1.jpg", "/a/IMG_2.jpg", "/a/IMG_1.cr2", "/a/x.png", "/b/IMG_1.jpg", "/b/y.jpg", "/a/.hidden",
             "/a/IMG_2.tar.gz", "/a/IMG_2.tar.xz"]
    groups = [[0, 3], [2, 5], [4, 6], [1, 7]]
    # 0 and 2 share (/a, IMG_1) -> groups 0 and 1 merge; /b/IMG_1.jpg is another directory
    assert scanner.merge_groups_by_stem(groups, paths) == [[0, 2, 3, 5], [1, 7], [4, 6]]
    # stems are "IMG_2.tar" for both archives: merged through the shared stem
    assert scanner.merge_groups_by_stem([[7, 0], [8, 4]], paths) == [[0, 4, 7, 8]]
    assert scanner.merge_groups_by_stem([[3, 1, 1]], paths) == [[1, 3]]
    assert scanner.merge_groups_by_stem([], paths) == []


def test_nccl_preload_keeps_torch_importable():
    """The library dlopens libnccl.so.2 at rh_group_create; a Python process that imports torch afterwards must
    still resolve torch's NCCL symbols (tools/fuzz_parity.py --multi found an ImportError when the older system
    NCCL got loaded first), so `_lib.Group` loads the pip-bundled copy first."""
    import subprocess
    import sys
    code = ("import sys\n"
            "from rupphash_b200 import _lib\n"
            "assert 'torch' not in sys.modules\n"
            "_lib._preload_bundled_nccl()\n"
            "import torch\n"
            "import torch.distributed\n"
            "print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
