"""GPU parity of the 64-bit DCT pHash (phash.rs:48-83) against the oracle's restatement.
The oracle itself is parity-unpinned for this path (the image / rustdct crates are not in the
reference tree); the device must match the oracle bit for bit."""
import numpy as np
import pytest

from rupphash_b200.synth import synth_images

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from rupphash_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("shape", [(384, 512, 3), (32, 32, 3), (32, 32), (100, 37, 4), (768, 1024, 3), (5, 9, 3),
                                   (64, 48), (333, 1000, 3)])
def test_phash_matches_oracle(ctx, orc, shape):
    from rupphash_b200 import phash
    h, w = shape[:2]
    ch = shape[2] if len(shape) == 3 else 1
    imgs = synth_images(5, h, w, seed=h + 3 * w, channels=ch)
    if ch == 1:
        imgs = imgs[..., 0]
    layout = {3: 0, 4: 1, 1: 2}[ch]
    hasher = phash.DctPhash(ctx)
    got, dih = hasher.hash_batch(imgs, want_dihedral=True)
    for k in range(len(imgs)):
        want, _ = orc.phash_image(imgs[k], layout=layout)
        assert int(got[k]) == want, (k, hex(int(got[k])), hex(want))
        assert [int(x) for x in dih[k]] == orc.phash_dihedral(want)
    assert hasher.hash_image(imgs[0]) == int(got[0])
    assert hasher.hash_image_invariant(imgs[0]) == orc.phash_rot_invariant(int(got[0]))


def test_phash_flat_and_extremes(ctx, orc):
    from rupphash_b200 import phash
    imgs = np.zeros((3, 64, 64, 3), np.uint8)
    imgs[1] = 255
    imgs[2, :, ::2] = 255
    got = phash.DctPhash(ctx).hash_batch(imgs)
    for k in range(3):
        assert int(got[k]) == orc.phash_image(imgs[k])[0]
