"""Multi-GPU plumbing: envs shard trivially over ranks (one process per GPU, contiguous global env
ids, no data-path collective); the only collective is the episode-statistics reduction."""
import os

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


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Restrict this process to the CPU cores of the NUMA node GPU ``device_index`` hangs off, so that the page-locked
    staging memory it allocates next is first-touched there and the per-step H2D / D2H copies of eight ranks do not all
    cross one socket's memory controller (SCALE_r01: end-to-end efficiency 0.48 at 8 GPUs with every rank on node 0).
    Never widens the affinity mask the process was started with (cgroup cpusets stay respected). Returns what was done."""
    out = {'bound': False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        ideal = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = sorted(ideal & allowed)
        out.update(ideal_cpus=len(ideal), allowed_cpus=len(allowed))
        if target and set(target) != allowed:
            os.sched_setaffinity(0, target)
            out.update(bound=True, cpus=f'{target[0]}-{target[-1]}', n_cpus=len(target))
        elif not target:
            out['note'] = 'the GPU-local cores are outside the allowed cpuset'
    except Exception as e:   # NVML missing, permission denied ...: report, never fail the run
        out['note'] = f'{type(e).__name__}: {e}'
    return out
