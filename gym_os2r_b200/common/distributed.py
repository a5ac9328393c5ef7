"""Multi-GPU plumbing: envs shard trivially over ranks (one process per GPU, contiguous global env
ids, no data-path collective); the only collective is the episode-statistics reduction."""
import torch
import torch.distributed as dist

STAT_KEYS = ('env_steps', 'episodes', 'done_task', 'done_timelimit', 'nonfinite_resets', 'sum_return', 'sum_length')


def shard_range(total_envs: int, rank: int, world_size: int):
    """Global env ids [first, first + count) owned by ``rank`` (remainder spread over the low ranks)."""
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def reduce_stats(stats: dict, device=None, group=None) -> dict:
    """Sum the per-rank episode statistics over all ranks (NCCL all-reduce of a 7-element vector on
    GPUs, gloo on CPU). Returns the global dict plus mean_return / mean_length."""
    vec = torch.tensor([float(stats[k]) for k in STAT_KEYS], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    out = {k: vec[i].item() for i, k in enumerate(STAT_KEYS)}
    ep = out['episodes']
    out['mean_return'] = out['sum_return'] / ep if ep > 0 else None
    out['mean_length'] = out['sum_length'] / ep if ep > 0 else None
    return out
