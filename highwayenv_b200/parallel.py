"""Multi-GPU sharding of independent envs (one process per GPU, `torch.distributed`).

The reference has no parallelism of its own; users vectorise with `SubprocVecEnv` /
`gym.vector` (reference scripts/sb3_highway_ppo.py:16-18, tests/envs/test_gym.py:158-164).
Here the envs of a batch are independent, so the batch is sharded by contiguous env-index
range: rank r of W owns global envs [r*E, (r+1)*E) and seeds them with `seed + global_index`,
which makes every env's trajectory independent of W.  There is NO collective on the step
path; the only optional exchange is the whole-batch observation gather below (NCCL over
NVLink on GPUs, gloo in the CPU tests) for a learner that wants all observations on every rank.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def env_range(rank: int, world_size: int, envs_per_rank: int) -> Tuple[int, int]:
    """Global env-index range [lo, hi) owned by `rank` (weak scaling: E envs on every rank)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    lo = rank * envs_per_rank
    return lo, lo + envs_per_rank


def split_envs(total_envs: int, world_size: int) -> Tuple[int, ...]:
    """Strong-scaling split of a fixed batch: near-even contiguous shards (first shards larger)."""
    q, r = divmod(total_envs, world_size)
    return tuple(q + 1 if k < r else q for k in range(world_size))


def all_gather_batch(local: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Concatenate every rank's `[E, ...]` tensor into `[W*E, ...]` in rank (= env index) order."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    local = local.contiguous()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                      device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out


def make_sharded(env_id: str, envs_per_rank: int, config=None, **kwargs):
    """`highwayenv_b200.make` for the calling rank (RANK / LOCAL_RANK / WORLD_SIZE from the env or an
    initialised process group): device cuda:LOCAL_RANK, env_index_offset = rank * envs_per_rank."""
    import os

    from . import make

    rank = dist.get_rank() if dist.is_initialized() else int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    lo, _ = env_range(rank, max(rank + 1, int(os.environ.get("WORLD_SIZE", "1"))), envs_per_rank)
    return make(env_id, num_envs=envs_per_rank, config=config, device=f"cuda:{local_rank}",
                env_index_offset=lo, **kwargs)
