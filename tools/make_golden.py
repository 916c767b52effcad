#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's shipped JPEG fixtures.

Runs only in the build container (needs /root/reference and Pillow).  PROVENANCE: the values
are produced by the CPU oracle (oracle/), NOT by the Rust reference -- the reference ships no
hash values for its images (tests/bench.jpg.txt etc. are licence notes) and cannot be compiled
here.  JPEG decode is Pillow/libjpeg-turbo, which may differ from zune-jpeg by +-1 LSB.

Per fixture image the file holds
    luma512     the oracle's luma plane after the Box pre-downsample (<= 512 px, u8): feeding it
                to the device as LUMA8 must reproduce the oracle's full-pipeline result
    hash/quality/coeffs/dihedral   oracle outputs for the full image
and for the two images wide enough, an RGB crop whose pre-downsample is exactly 2x
    crop_rgb + crop_hash/crop_quality/crop_coeffs/crop_dihedral
and prophecy1 also carries full_rgb, the whole decoded image (general Box pre-downsample path).
"""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

REF = "/root/reference/tests"
OUT = os.path.join(ROOT, "tests", "golden")
FIXTURES = {
    "bench": ("bench.jpg", (1024, 768)),
    "kaanapali": ("Kaanapali_beach_sunrise_on_Maui_Hawaii.720p.jpg", (1024, 720)),
    "prophecy1": ("Prophecy_Has_Been_Fulfilled_1.jpg", None),
    "prophecy2": ("Prophecy_Has_Been_Fulfilled_2.jpg", None),
}


def full_pipeline(rgb):
    h, w = rgb.shape[:2]
    luma = oracle.luma601(rgb).reshape(h, w)
    if w > 512 or h > 512:
        nw, nh = oracle.target_dimensions(w, h)
        luma = oracle.resize_box_u8(luma, nw, nh)
    coeffs, q, _ = oracle.pdq_from_luma(luma)
    return luma, coeffs, q


def main():
    oracle.build()
    os.makedirs(OUT, exist_ok=True)
    for name, (fn, crop) in FIXTURES.items():
        rgb = np.asarray(Image.open(os.path.join(REF, fn)).convert("RGB"))
        luma, coeffs, q = full_pipeline(rgb)
        ref = oracle.pdq_features(rgb)
        assert ref is not None and np.array_equal(ref[0], coeffs) and ref[1] == q
        d = dict(luma512=luma, hash=oracle.to_hash(coeffs), quality=np.float32(q), coeffs=coeffs,
                 dihedral=oracle.dihedral(coeffs), src_shape=np.array(rgb.shape))
        if crop:
            cw, ch = crop
            h, w = rgb.shape[:2]
            y0, x0 = (h - ch) // 2, (w - cw) // 2
            c = np.ascontiguousarray(rgb[y0:y0 + ch, x0:x0 + cw])
            cc, cq, _ = oracle.pdq_features(c)
            d.update(crop_rgb=c, crop_hash=oracle.to_hash(cc), crop_quality=np.float32(cq), crop_coeffs=cc,
                     crop_dihedral=oracle.dihedral(cc))
        if name == "prophecy1":
            # the whole decoded image (780 x 768 -> 512 x 504, a non-2x Box pre-downsample): the
            # device must reproduce hash / quality / coefficients above from these pixels
            d["full_rgb"] = rgb
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **d)
        print(name, rgb.shape, "->", luma.shape, bytes(d["hash"]).hex(), float(q), os.path.getsize(path))


if __name__ == "__main__":
    main()
