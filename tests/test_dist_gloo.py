"""CPU, world_size 2 over gloo: the host logic of scanner.group_files_sharded (forest all-gather,
merge, edge-count all-reduce).  The per-rank tile kernel and the merge are stood in for by the
oracle's rank functions (there is no GPU here); on a GPU box the same function runs the device
calls and NCCL."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    import oracle
    from rupphash_b200 import scanner
    from rupphash_b200.synth import planted_hashes, random_variants

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hashes, low_conf = planted_hashes(3000, seed=17)
    variants = random_variants(hashes, seed=2)
    shard = lambda: oracle.group_tiles_rank(hashes, 31, 256, rank, world, variants=variants, low_conf=low_conf)
    labels, total = scanner.group_files_sharded(hashes, 31, variants=variants, low_conf=low_conf, shard_fn=shard,
                                                merge_fn=oracle.merge_parents)
    np.save(os.path.join(out_dir, f"labels{rank}.npy"), np.asarray(labels))
    np.save(os.path.join(out_dir, f"count{rank}.npy"), np.array([total]))
    dist.destroy_process_group()


def test_sharded_grouping_world2(tmp_path, orc):
    from rupphash_b200.synth import planted_hashes, random_variants
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    hashes, low_conf = planted_hashes(3000, seed=17)
    variants = random_variants(hashes, seed=2)
    ref_labels, ref_cnt, _ = orc.group_generic(hashes, 31, variants=variants, low_conf=low_conf, use_mih=True)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"labels{r}.npy"), ref_labels)
        assert int(np.load(tmp_path / f"count{r}.npy")[0]) == ref_cnt
