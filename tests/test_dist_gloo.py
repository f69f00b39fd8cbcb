"""CPU tier for the N>1 path: env sharding and the episode-statistics all-reduce (gloo, world_size 2).

The step path has no collective (episodes are independent); the only exchange is the optional
all-reduce(sum) of the 8-double statistics vector, which runs over NCCL on GPUs and is exercised
here over gloo with the same code (roborugby_b200/stats.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from roborugby_b200.stats import allreduce_stats, shard_envs, summarize


def test_shard_envs_partitions_exactly():
    for total, world in [(65536, 1), (65536, 8), (1048576, 8), (1000, 3), (7, 8)]:
        spans = [shard_envs(total, r, world) for r in range(world)]
        assert sum(n for n, _ in spans) == total
        off = 0
        for n, o in spans:
            assert o == off
            off += n
        assert max(n for n, _ in spans) - min(n for n, _ in spans) <= 1
    with pytest.raises(ValueError):
        shard_envs(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local, offset = shard_envs(1001, rank, world)
    # episodes, return_h, return_g, length, naughty, errors, steps, reserved
    local = torch.tensor([n_local, 2.0 * (rank + 1), -1.0 * (rank + 1), 300.0 * n_local, rank, 0, 16.0 * n_local, 0],
                         dtype=torch.float64)
    before = local.clone()
    glob = allreduce_stats(local)
    assert torch.equal(local, before), "the per-rank vector must not be modified"
    q.put((rank, n_local, offset, glob))
    dist.barrier()
    dist.destroy_process_group()


def test_stats_allreduce_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, n0, o0, g0), (r1, n1, o1, g1) = out
    assert (n0, o0, n1, o1) == (501, 0, 500, 501)
    assert g0 == g1, "every rank sees the same global statistics"
    assert g0["episodes"] == 1001 and g0["return_happy"] == 6.0 and g0["return_grumpy"] == -3.0
    assert g0["steps"] == 16 * 1001 and g0["naughty"] == 1
    assert g0["mean_length"] == 300.0
    assert g0["mean_return_happy"] == pytest.approx(6.0 / 1001)


def test_summarize_handles_zero_episodes():
    d = summarize([0, 0, 0, 0, 0, 0, 5, 0])
    assert d["mean_return_happy"] == 0.0 and d["steps"] == 5
