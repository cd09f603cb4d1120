"""CPU tier, world_size 2 over gloo: the host-side logic of the multi-GPU path -- env
sharding and the single statistics all-reduce (sum / min / max semantics)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from powergridworld_b200.multiagent_env import reduce_stats, shard_envs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = shard_envs(total, rank, world)
    # a rank's local statistics as pgw_stats would produce them for its env block
    envs = np.arange(first, first + n, dtype=np.float64)
    local = torch.tensor([n * 10.0, envs.sum(), 2 * envs.sum(), 0.5 * n, float(rank), 7.0 * n,
                          0.9 + 0.01 * rank, 1.0 + 0.01 * rank], dtype=torch.float64)
    red = reduce_stats(local)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([[first, n], red.numpy()]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_stats_allreduce_world2(tmp_path):
    world, total = 2, 4097                       # ragged split
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert (r0[0], r0[1]) == (0, 2049) and (r1[0], r1[1]) == (2049, 2048)
    np.testing.assert_array_equal(r0[2:], r1[2:])            # every rank holds the same result
    e = np.arange(total, dtype=np.float64)
    want = [total * 10.0, e.sum(), 2 * e.sum(), 0.5 * total, 1.0, 7.0 * total, 0.9, 1.01]
    np.testing.assert_allclose(r0[2:], want, rtol=1e-12)


def test_shard_envs_covers_everything():
    for total, world in [(4096, 8), (1000003, 8), (5, 8), (16384, 4)]:
        blocks = [shard_envs(total, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and sum(n for _, n in blocks) == total
        for (f0, n0), (f1, _) in zip(blocks, blocks[1:]):
            assert f0 + n0 == f1


def test_reduce_stats_is_identity_without_process_group():
    s = torch.arange(8, dtype=torch.float64)
    assert reduce_stats(s) is s
