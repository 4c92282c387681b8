"""Multi-GPU layout: environments are independent, so they shard by contiguous global index,
one process per GPU (torchrun), with NO collective on the simulation path.  torch.distributed is
used only by callers that need a cross-rank reduction (bench.py: max of the device time, sums of
env-steps; a trainer: gradient / rollout-statistics all-reduce).

Per-env random streams (scene sampling) are keyed by the GLOBAL env id (``env_id_offset + i``),
so a run on G GPUs reproduces the 1-GPU run env for env (tests/test_gpu_parity.py
test_shard_invariance; tests/test_distributed_cpu.py over gloo).
"""
from __future__ import annotations

import os


def world_info():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_for_rank(total_envs: int, rank: int, world_size: int):
    """Contiguous shard [offset, offset + count) of `total_envs` for `rank`; the first
    total_envs % world_size ranks get one env more."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def make_shard(total_envs: int, params=None, seed: int = 0, *, rank=None, world_size=None, device=None, _cdll=None):
    """The SalpBatch of this rank's shard of a `total_envs`-env job."""
    from .batch import SalpBatch
    r, lr, w = world_info()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    device = lr if device is None else device
    offset, count = shard_for_rank(total_envs, rank, world_size)
    return SalpBatch(count, params, seed=seed, env_id_offset=offset, device=device, _cdll=_cdll)


def allreduce_scalar(value: float, op: str = "max", device=None) -> float:
    """max / sum of a python float over all ranks (identity when not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())
