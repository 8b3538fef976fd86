"""N > 1 host path on CPU: two processes over gloo shard a ragged batch, each solves its slice (with the CPU
oracle standing in for the GPU solver -- this test is about partitioning and the u0 all-gather, not the
kernels) and the gathered u0 equals the unsharded solve."""
import os
import socket

import numpy as np
import pytest

from crazyflie_nmpc_b200 import sharding, workloads as wl

TS = 0.015


def test_shard_bounds_cover_batch_exactly():
    for B in (1, 5, 8, 65536, 262144 + 3):
        for world in (1, 2, 3, 8):
            cuts = [sharding.shard_bounds(B, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def _worker(rank, world, port_no, B, N, q):
    import torch
    import torch.distributed as dist
    from oracle.oracle import Port
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = sharding.shard_workload(wl.helix_batch(B, N, seed=9), world, rank)
    x, u = w["x_init"].copy(), w["u_init"].copy()
    Port().batch(N, TS, w["x0"], w["yref"], w["yref_e"], x, u)
    u0_all = sharding.gather_u0(torch.from_numpy(u[:, 0].copy()), B)
    if rank == 0:
        q.put(u0_all.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_solve_gathers_u0(port):
    import torch.multiprocessing as mp
    B, N, world = 5, 10, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, B, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    u0_all = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = wl.helix_batch(B, N, seed=9)
    x, u = w["x_init"].copy(), w["u_init"].copy()
    port.batch(N, TS, w["x0"], w["yref"], w["yref_e"], x, u)
    assert u0_all.shape == (B, 4) and np.array_equal(u0_all, u[:, 0])
