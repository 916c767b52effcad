"""The one-process-per-GPU route on real GPUs: N ranks over NCCL (spawned like torchrun spawns them), each rank
runs its share of the tiles on its own GPU (rh_hamming_group_shard), forests all-gathered over NVLink, merged;
labels and comparison_count on every rank equal the CPU oracle.  Needs >= 2 GPUs (gpurun --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, similarity):
    import torch
    import torch.distributed as dist

    from rupphash_b200 import _lib, scanner
    from rupphash_b200.synth import planted_hashes, random_variants

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = _lib.Context(rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    hashes, low_conf = planted_hashes(80_000, seed=23)
    variants = random_variants(hashes, seed=4)
    d_h, d_l, d_v = (torch.from_numpy(x).cuda() for x in (hashes, low_conf, variants))
    labels, total = scanner.group_files_sharded(d_h, similarity, variants=d_v, low_conf=d_l, ctx=ctx)
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"labels{rank}.npy"), labels.cpu().numpy().view(np.uint32))
    np.save(os.path.join(out_dir, f"count{rank}.npy"), np.array([total]))
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("similarity", [31, 40])
def test_sharded_grouping_over_nccl(tmp_path, orc, similarity):
    import torch
    from rupphash_b200.synth import planted_hashes, random_variants
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus N)")
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), similarity), nprocs=world, join=True)
    hashes, low_conf = planted_hashes(80_000, seed=23)
    variants = random_variants(hashes, seed=4)
    ref_labels, ref_cnt, _ = orc.group_generic(hashes, similarity, variants=variants, low_conf=low_conf, threads=8)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"labels{r}.npy"), ref_labels), r
        assert int(np.load(tmp_path / f"count{r}.npy")[0]) == ref_cnt
