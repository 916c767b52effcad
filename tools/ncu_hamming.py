#!/usr/bin/env python
"""Fixed workload for an ncu capture of the Hamming tile kernel at the bench size: 500k planted hashes, similarity 31."""
import sys
import torch
sys.path.insert(0, '/root/repo')
from rupphash_b200 import _lib, scanner
from rupphash_b200.synth import planted_hashes
ctx = _lib.Context(0)
hashes, low_conf = planted_hashes(500_000, seed=0xB200, n_clusters=5000, identical_block=1000, threshold=31)
d_h, d_l = torch.from_numpy(hashes).cuda(), torch.from_numpy(low_conf).cuda()
for _ in range(2):
    labels, cnt = scanner.group_labels(d_h, 31, low_conf=d_l, ctx=ctx)
print("ok", cnt, ctx.last_kernel_time())
