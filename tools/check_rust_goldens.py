#!/usr/bin/env python
"""Compare the oracle with vectors produced by the real reference (tools/gen_goldens.rs, see
tools/regen_with_rust.md).  usage: check_rust_goldens.py tests/golden/rust/goldens_rust.jsonl"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def main():
    path = sys.argv[1]
    base = os.path.dirname(path)
    oracle.build()
    bad = 0
    for line in open(path):
        g = json.loads(line)
        rgb = np.fromfile(os.path.join(base, os.path.basename(g["file"]) + ".rgb"), np.uint8).reshape(g["h"], g["w"], 3)
        assert hashlib.sha256(rgb.tobytes()).hexdigest() == g["rgb_sha256"], "pixel file does not match the vector"
        coeffs, q, _ = oracle.pdq_features(rgb)
        h = oracle.to_hash(coeffs)
        want_h = np.frombuffer(bytes.fromhex(g["pdq_hash"]), np.uint8)
        dist = int(np.unpackbits(h ^ want_h).sum())
        want_c = np.array(g["pdq_coeffs"], np.uint32)
        cdiff = int((coeffs.reshape(256).view(np.uint32) != want_c).sum())
        qdiff = abs(float(np.float32(q)) - float(np.array([g["pdq_quality"]], np.uint32).view(np.float32)[0]))
        dih = oracle.dihedral(coeffs)
        ddist = [int(np.unpackbits(dih[k] ^ np.frombuffer(bytes.fromhex(x), np.uint8)).sum()) for k, x in enumerate(g["pdq_dihedral"])]
        res = "n/a"
        if g.get("resized"):
            luma = oracle.luma601(rgb)
            r = oracle.resize_box_u8(luma, g["resized"]["w"], g["resized"]["h"])
            res = "same" if hashlib.sha256(r.tobytes()).hexdigest() == g["resized"]["luma_sha256"] else "DIFFERENT"
        ph, _ = oracle.phash_image(rgb)
        pdist = bin(int(ph) ^ int(g["phash"], 16)).count("1")
        print(f"{g['file']}: pdq distance {dist}, quality diff {qdiff:.3g}, coefficient words differing {cdiff}, "
              f"dihedral distances {ddist}, resized luma {res}, phash distance {pdist}")
        bad += dist + cdiff + sum(ddist) + pdist + (res == "DIFFERENT")
    print("ORACLE PINNED" if bad == 0 else "ORACLE DIFFERS FROM THE REFERENCE: see the lines above")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
