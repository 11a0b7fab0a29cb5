"""Multi-GPU drivers, one process per GPU.

1. Independent samples: dealt to ranks with no data-path collective; `torch.distributed` only carries the small result
   records back to rank 0 (shard_indices / merge_shards / run_sharded).
2. One large diploid DP, row-sharded (north_star: "h1-row tiles of the diploid matrix"): RowShardedDip.  The exchange is
   fused into the sweep kernel (rows and barrier arrivals go through NVLink peer mappings); `torch.distributed` only
   all-gathers the CUDA IPC handles once and provides the host barrier before each launch.

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


class RowShardedDip:
    """One diploid DP over all the GPUs of the process group (include/dipgenie_cuda.h: dg_dip_create_sharded).

    Every rank passes the same LevelGraph.  Wide level transitions (approximator.cpp:627-701) are split by destination
    row over world x ctas CTAs; narrow ones run redundantly on every rank; every rank ends with the full result.

        prob = RowShardedDip(ctx, graph, R, dist)      # collective: all ranks
        out = prob.run()                               # collective; the same dict on every rank
    """

    def __init__(self, ctx, graph, R: int, dist, ctas: int = 0):
        if dist is None or not dist.is_initialized():
            raise RuntimeError("RowShardedDip needs an initialised torch.distributed process group")
        self.dist = dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.problem = ctx.dip_create_sharded(graph, R, self.rank, self.world, ctas)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.problem.ipc_export())
        self.problem.ipc_attach(handles)

    def run(self, profile: bool = False) -> dict:
        self.problem.shard_arm()
        self.dist.barrier()            # every rank's counters are reset before any sweep can arrive on them
        self.problem.run(profile=profile)
        out = self.problem.result()
        self.dist.barrier()            # nobody re-arms while a peer's kernel may still be running
        return out

    def stats(self) -> dict:
        return self.problem.stats()

    def close(self):
        self.dist.barrier()
        self.problem.close()
