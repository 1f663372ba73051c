"""World-size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding and the flat-bucket gradient
all-reduce used by the training step."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wind_downscaling_gan_b200.train.dist import Comm, shard_batch
    comm = Comm()
    assert (comm.world, comm.rank) == (world, rank)
    grads = {"a": torch.full((3, 4), float(rank + 1)), "bn/gamma": torch.full((5,), 10.0 * (rank + 1)),
             "c": torch.arange(6, dtype=torch.float32) * (rank + 1)}
    comm.allreduce_grads(grads, skip=["bn/gamma"])
    s = torch.ones(2, 3) * (rank + 1)
    comm.allreduce_sum(s)
    q.put((rank, grads["a"].clone(), grads["bn/gamma"].clone(), grads["c"].clone(), s.clone(), shard_batch(7, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_bucket_and_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, a, gamma, c, s, shard in res:
        assert torch.equal(a, torch.full((3, 4), 3.0))                       # 1 + 2
        assert torch.equal(gamma, torch.full((5,), 10.0 * (rank + 1)))        # skipped: already global
        assert torch.equal(c, torch.arange(6, dtype=torch.float32) * 3)
        assert torch.equal(s, torch.full((2, 3), 3.0))
    assert [r[5] for r in res] == [(0, 4), (4, 7)]


def test_single_process_comm_is_noop():
    from wind_downscaling_gan_b200.train.dist import Comm, shard_batch
    c = Comm()
    assert c.world == 1 and c.rank == 0
    g = {"x": torch.ones(3)}
    assert c.allreduce_grads(g)["x"].sum() == 3
    assert shard_batch(8, 0, 1) == (0, 8) and shard_batch(8, 3, 4) == (6, 8)
