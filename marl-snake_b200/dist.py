"""Multi-GPU plumbing: environments are independent, so a job shards contiguous env ranges over
ranks (one process per GPU) with NO collective on the step path.  The only exchange is an optional
end-of-rollout all-reduce of the 8-double statistics vector (NCCL on GPUs, gloo in CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total_envs, rank, world_size):
    """Contiguous [start, stop) of global env ids owned by `rank`; sizes differ by at most one.
    Philox streams are keyed by GLOBAL env id (SnakeBatch(env_id_offset=start)), so the union of the
    shards reproduces the single-device run bit for bit."""
    base, rem = divmod(int(total_envs), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_stats(stats, group=None):
    """In-place SUM all-reduce of a statistics vector (torch tensor, any device) across ranks.
    No-op when torch.distributed is not initialised (single GPU)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats
