"""N>1 host logic on CPU: world_size-2 gloo.  Each rank 'samples' its shard (here with
the ORACLE standing in for the GPU sampler — this is a test of sharding + gather, not of
the product path), the draws are all-gathered, and the result must equal the
single-process run in global chain order."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, total, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from __graft_entry__ import load_package
    import pyoracle as po
    pkg = load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, cnt = pkg.parallel.shard(total, world, rank)
    D = 6
    local = np.stack([po.sample_tree(0, po.random_position(9, off + k, D), 0.5, 9, off + k, 0)["q"]
                      for k in range(cnt)])
    full = pkg.parallel.gather_draws(torch.from_numpy(local), total)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_two_rank_shard_and_gather(tmp_path, po, total):
    out = str(tmp_path / "g.npy")
    port = 29500 + (os.getpid() + total) % 2000
    mp.spawn(_worker, args=(2, total, port, out), nprocs=2, join=True)
    got = np.load(out)
    ref = np.stack([po.sample_tree(0, po.random_position(9, k, 6), 0.5, 9, k, 0)["q"] for k in range(total)])
    assert np.array_equal(got, ref)


def test_shard_partition(pkg):
    for total in (1, 7, 8, 65536, 524288):
        for world in (1, 2, 3, 8):
            blocks = [pkg.parallel.shard(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (o1, c1), (o2, _) in zip(blocks, blocks[1:]):
                assert o1 + c1 == o2
    with pytest.raises(ValueError):
        pkg.parallel.shard(8, 2, 2)
