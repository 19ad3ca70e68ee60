"""world_size-2 `gloo` tests (CPU): the host-side logic of the sharded path — contiguous shards, the packed
(H, b, sum) all-reduce and its replication on every rank — with the oracle standing in for the device pass.
Mirrors tst/multiple_objectives.cpp:112-125 (split cost == single cost) as a distributed sum."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from moptimizer_0_b200 import sharding
from oracle import oracle_py as orc
from tests.common import fachada, rel_err


def test_shard_ranges_partition_everything():
    for n in (0, 1, 7, 100, 29310, 10**9 + 7):
        for w in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    for P in (1, 2, 6, 15):
        A = rng.normal(size=(P, P))
        H = A + A.T
        b = rng.normal(size=P)
        v = sharding.pack(H, b, 3.5)
        assert v.shape[0] == sharding.packed_size(P)
        H2, b2, s2 = sharding.unpack(v, P)
        assert np.array_equal(H, H2) and np.array_equal(b, b2) and s2 == 3.5


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    src, tgt, _, _ = fachada()
    n = src.shape[0]
    lo, hi = sharding.shard_range(n, rank, world)
    x = [0.5, -0.2, 0.1, 0.05, 0.02, -0.03]
    cost = orc.Cost(orc.P2P, 6, 3, hi - lo, a=src[lo:hi], b=tgt[lo:hi], jac_mode=orc.JAC_ANALYTICAL,
                    loss=orc.LOSS_HUBER, loss_param=5.0)
    H, b, s = orc.linearize(cost, x)
    v = torch.from_numpy(sharding.pack(H, b, s))
    dist.all_reduce(v, op=dist.ReduceOp.SUM)  # the one collective of the path
    # every rank must hold bit-identical reduced values (=> identical accept/reject decisions)
    gathered = [torch.zeros_like(v) for _ in range(world)]
    dist.all_gather(gathered, v)
    identical = all(torch.equal(g, gathered[0]) for g in gathered)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.save(out, np.concatenate([v.numpy(), [float(identical), t.item()]]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_sum_equals_single(tmp_path, world):
    out = str(tmp_path / "reduced.npy")
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    res = np.load(out)
    v, identical, tmax = res[:-2], res[-2], res[-1]
    assert identical == 1.0 and tmax == float(world)
    src, tgt, _, _ = fachada()
    x = [0.5, -0.2, 0.1, 0.05, 0.02, -0.03]
    cost = orc.Cost(orc.P2P, 6, 3, src.shape[0], a=src, b=tgt, jac_mode=orc.JAC_ANALYTICAL, loss=orc.LOSS_HUBER,
                    loss_param=5.0)
    H1, b1, s1 = orc.linearize(cost, x)
    H, b, s = sharding.unpack(v, 6)
    # summation order differs from the single-rank run: equal to rounding, not bitwise
    assert rel_err(H, H1) < 1e-12 and rel_err(b, b1) < 1e-12 and abs(s - s1) <= 1e-12 * s1
