"""World-size-2 gloo test of the multi-GPU path's host logic (runs on CPU): evolving disjoint blocks of the points
on two ranks and all-gathering them reproduces the single-process evolve bit for bit (points are independent:
gple/evolve.cpp:392-420), for even and uneven partitions."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussian_process_liouville_equation_b200 import sharding
from gaussian_process_liouville_equation_b200 import synthetic as syn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ORACLE_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc

    g = syn.rng(50, 0)
    r = np.stack([-1.5 + syn.SIGMA_X * g.standard_normal(total), syn.P0 + syn.SIGMA_P * g.standard_normal(total)], 1)
    an = np.array([-1.5, syn.P0, syn.SIGMA_X, syn.SIGMA_P, 1.0, 0.0, 0.0, 0.0])
    pts = syn.points_aos(r, orc.initial_distribution(an, r, 0, 0))
    lo, hi = sharding.partition(total, rank, world)
    mine, _, _ = orc.evolve(1, pts[lo:hi], None, None, syn.MASS, 2.0, analytic=an)
    full = sharding.all_gather_points(torch.from_numpy(mine), total).numpy()
    if rank == 0:
        whole, _, _ = orc.evolve(1, pts, None, None, syn.MASS, 2.0, analytic=an)
        ret["equal"] = bool(np.array_equal(full, whole))
    dist.destroy_process_group()


def _mc_worker(rank, world, port, total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ORACLE_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc

    g = syn.rng(51, 0)
    pts = np.zeros((total, 4))
    pts[:, 0] = syn.X0 + syn.SIGMA_X * g.standard_normal(total)
    pts[:, 1] = syn.P0 + syn.SIGMA_P * g.standard_normal(total)
    an = [syn.X0, syn.P0, syn.SIGMA_X, syn.SIGMA_P, 0.8, 0.6, 0.0, 0.4]
    lo, hi = sharding.partition(total, rank, world)
    # every rank walks its block of the chains on the streams those chains own in the whole set (chain0 = lo)
    mine, _, _ = orc.markov_chains(pts[lo:hi], 40, 0.5, 23, 2, 1, 0, analytic=an, chain0=lo)
    full = sharding.all_gather_points(torch.from_numpy(mine), total).numpy()
    if rank == 0:
        whole, _, _ = orc.markov_chains(pts, 40, 0.5, 23, 2, 1, 0, analytic=an)
        ret["equal"] = bool(np.array_equal(full, whole))
    dist.destroy_process_group()


def test_sharded_markov_chains_plus_allgather_equal_single_process():
    """Metropolis chains are keyed by (seed, stream, chain index): a block of chains gives the same result on any rank."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_mc_worker, args=(2, _free_port(), 37, ret), nprocs=2, join=True)
    assert ret["equal"]


@pytest.mark.parametrize("total", [64, 37])
def test_sharded_evolve_plus_allgather_equals_single_process(total):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), total, ret), nprocs=2, join=True)
    assert ret["equal"]


def test_partition_covers_everything():
    for total in (0, 1, 7, 100000):
        for world in (1, 2, 3, 8):
            blocks = [sharding.partition(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(sharding.counts(total, world)) - min(sharding.counts(total, world)) <= 1
