#!/usr/bin/env python
"""configs[4] parity at FULL size: 1M files x 8 real dihedral variants grouped by a plain C++ program (no Python,
no torch in the process that drives the GPUs: tests/cpp/group_multi_harness.cpp over the C ABI) on every GPU of the
box, against the CPU oracle's MIH search of the same input.  Writes gpurun_out/config5_check.json.
usage: config5_check.py [n_files] [similarity]"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import oracle  # noqa: E402
from rupphash_b200 import _lib  # noqa: E402
from test_gpu_group_multi import run_harness  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    sim = int(sys.argv[2]) if len(sys.argv) > 2 else 31
    ctx = _lib.Context(0)
    d_h, d_v, d_l = bench.dihedral_coeff_dataset(torch, ctx, n, seed=0x5EED)
    h, v, lc = d_h.cpu().numpy(), d_v.cpu().numpy(), d_l.cpu().numpy()
    del d_h, d_v, d_l
    ctx.close()
    torch.cuda.empty_cache()
    oracle.build()
    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as tmp:
        res = run_harness(tmp, h, sim, variants=v, low_conf=lc)
    t0 = time.perf_counter()
    want, want_cnt, _ = oracle.group_generic(h, sim, variants=v, low_conf=lc, threads=cores)
    cpu_s = time.perf_counter() - t0
    out = {"n_files": n, "query_rows": 8 * n, "similarity": sim, "n_gpus": int(res["n_gpus"]), "nccl": int(res["nccl"]),
           "group_wall_ms": res["wall_ms"], "tile_kernel_ms_slowest_gpu": res["tile_ms_max"],
           "tile_kernel_gpu_ms_sum": res["tile_ms_sum"], "edges": int(res["edges"]),
           "cpu_oracle_s": cpu_s, "cpu_cores": cores,
           "multi_gpu_labels_identical_to_oracle": bool(np.array_equal(res["multi"], want)),
           "single_gpu_labels_identical_to_oracle": bool(np.array_equal(res["single"], want)),
           "edges_identical_to_oracle": bool(int(res["edges"]) == want_cnt),
           "groups": int((np.bincount(want, minlength=n) > 1).sum()), "harness": res["stdout"].strip()}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"config5_check_{out['n_gpus']}gpu_sim{sim}.json"), "w"), indent=1)
    print(json.dumps(out))
    assert out["multi_gpu_labels_identical_to_oracle"] and out["edges_identical_to_oracle"]


if __name__ == "__main__":
    main()
