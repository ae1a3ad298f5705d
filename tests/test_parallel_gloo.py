"""world_size-2 gloo tests of the multi-GPU host path (runs on CPU): index sharding, the single
all-reduce of rollout statistics and the grid-search cost-vector assembly."""

import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from data_driven_quad_control_b200.parallel import (
        allreduce_cost_vector,
        allreduce_rollout_stats,
        shard_range,
        shard_round_robin,
    )

    lo, hi = shard_range(1001, rank, world)
    stats = {"env_steps": float(hi - lo) * 10, "resets": float(rank + 1), "reward_sum": -0.5 * (rank + 1)}
    total = allreduce_rollout_stats(stats)
    owned = shard_round_robin(9, rank, world)
    costs = torch.tensor([float(i) ** 2 for i in owned])
    full = allreduce_cost_vector(costs, owned, 9)
    out.put((rank, total, full.tolist(), (lo, hi)))
    dist.destroy_process_group()


def test_stats_allreduce_and_cost_vector_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, total, full, (lo, hi) in results:
        assert total == {"env_steps": 10010.0, "resets": 3.0, "reward_sum": -1.5}
        assert full == [float(i) ** 2 for i in range(9)]
    assert sorted(r[3] for r in results) == [(0, 501), (501, 1001)]


def test_allreduce_is_identity_without_process_group():
    sys.path.insert(0, ROOT)
    from data_driven_quad_control_b200.parallel import allreduce_rollout_stats

    assert allreduce_rollout_stats({"a": 1.5}) == {"a": 1.5}
