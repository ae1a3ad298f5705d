"""Multi-GPU plumbing for the hot path: one process per GPU, environments / controllers sharded
by index, and ONE collective -- the final all-reduce of rollout statistics and grid-search cost
vectors.  There is no per-step exchange: each rank owns its own `HoverEnv`
instance, so the reference's two batch-global quirks (whole-batch PID reset, all-env rel_pos
refresh) stay per instance, exactly as in a reference run with num_envs/W envs per process.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous index range [lo, hi) owned by `rank` when `n_total` units are split as evenly
    as possible over `world` ranks (the first n_total % world ranks get one extra unit)."""
    if world < 1 or not 0 <= rank < world or n_total < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_world() -> tuple[int, int]:
    """(rank, world size) of the current process group; (0, 1) without one."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_round_robin(n_total: int, rank: int, world: int) -> list[int]:
    """Indices owned by `rank` under round-robin assignment (used for grid-search combinations,
    whose cost depends on the (n, L) shape bucket, to balance the buckets over ranks)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad shard arguments")
    return list(range(rank, n_total, world))


def _device_for_backend() -> torch.device:
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def allreduce_rollout_stats(stats: dict[str, float]) -> dict[str, float]:
    """Sum a dict of rollout / episode statistics over all ranks (float64, one all-reduce).
    Without an initialised process group it returns the input unchanged."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(stats)
    keys = sorted(stats)
    buf = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=_device_for_backend())
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(keys, buf.cpu().tolist())}


def allreduce_cost_vector(local_costs: torch.Tensor, owned: list[int], n_total: int) -> torch.Tensor:
    """Assemble the full grid-search cost vector: each rank contributes the costs of the
    combinations it owns (`owned` indices), everything else zero, and the vectors are summed."""
    full = torch.zeros((n_total,), dtype=torch.float64, device=local_costs.device)
    if len(owned):
        full[torch.as_tensor(owned, device=local_costs.device)] = local_costs.to(torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl" and full.device.type != "cuda":
            full = full.cuda()
        dist.all_reduce(full, op=dist.ReduceOp.SUM)
    return full
