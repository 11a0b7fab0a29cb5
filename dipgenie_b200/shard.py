"""Multi-GPU driver for independent samples: one process per GPU, samples dealt to ranks with no
data-path collective; `torch.distributed` only carries the small result records back to rank 0.

The reference runs one DipGenie process per sample (data/run_DipGenie_batch.sh:21-39: the 22-sample
leave-one-out study); here a rank takes its share of the samples and runs them side by side on its GPU with
`Context.dp_diploid_batch` (include/dipgenie_cuda.h: dg_dp_diploid_batch)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence


def shard_indices(n_samples: int, world: int, rank: int) -> List[int]:
    """Samples of `rank`: round-robin deal (sample i -> rank i % world), so that a batch sorted by size stays
    balanced.  Every sample belongs to exactly one rank."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    return list(range(rank, n_samples, world))


def merge_shards(n_samples: int, shards: Sequence[Sequence], world: int) -> list:
    """Inverse of shard_indices: shards[r][j] is the result of sample r + j * world."""
    out: list = [None] * n_samples
    for r in range(world):
        idx = shard_indices(n_samples, world, r)
        if len(shards[r]) != len(idx):
            raise ValueError(f"rank {r} returned {len(shards[r])} results for {len(idx)} samples")
        for j, i in enumerate(idx):
            out[i] = shards[r][j]
    return out


def run_sharded(samples: Sequence, run_local: Callable[[list], list], dist=None) -> Optional[list]:
    """Deal `samples` to the ranks of the default process group, run `run_local` on the local share (on a GPU
    box: lambda s: ctx.dp_diploid_batch(s, R)), and return all results in sample order on rank 0 (None elsewhere).
    With no process group it is a plain local call."""
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return run_local(list(samples))
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = [samples[i] for i in shard_indices(len(samples), world, rank)]
    local = run_local(mine)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)      # result records only: a few hundred bytes per sample
    if rank != 0:
        return None
    return merge_shards(len(samples), gathered, world)
