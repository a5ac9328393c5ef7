"""CPU suite, world_size = 2 over gloo: the N>1 path's host logic — env sharding by global id (results
independent of the rank count) and the episode-statistics reduction (the path's only collective)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total_envs, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import oracle
    from gym_os2r_b200.common.distributed import reduce_stats, shard_range
    from helpers import make_config
    task, cm, cfg = make_config('fixed_hip', reward='BalancingV1', reset_randomized=True, randomize_params=True,
                                randomize_gravity=True, auto_reset=True, max_episode_steps=4)
    first, count = shard_range(total_envs, rank, world)
    # The CPU oracle stands in for the GPU shard here (no device in this container): the sharding rule and the
    # RNG keying by GLOBAL env id are what is under test.
    shard = oracle.Oracle(cm.struct, cfg, count, first_env_id=first, seed=5)
    shard.reset()
    rng = np.random.RandomState(0)
    acts = rng.uniform(-1, 1, (9, total_envs, 2))
    episodes = 0
    ret_sum = 0.0
    for t in range(9):
        obs, rew, done, term, info = shard.step(acts[t, first:first + count])
        episodes += int(done.sum())
    stats = {'env_steps': 9 * count, 'episodes': episodes, 'done_task': 0, 'done_timelimit': episodes,
             'nonfinite_resets': 0, 'sum_return': float(rank + 1), 'sum_length': 4.0 * episodes}
    red = reduce_stats(stats)
    np.savez(os.path.join(out_dir, f'rank{rank}.npz'), state=shard.state, params=shard.params, obs=obs,
             first=first, count=count, red=np.array([red[k] for k in ('env_steps', 'episodes', 'sum_return', 'sum_length')]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats_reduction(tmp_path):
    total = 37                      # odd on purpose: ragged shards (19 + 18)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / 'rank0.npz'), np.load(tmp_path / 'rank1.npz')
    assert (int(r0['first']), int(r0['count'])) == (0, 19) and (int(r1['first']), int(r1['count'])) == (19, 18)
    # single-process run over the union
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import oracle
    from helpers import make_config
    task, cm, cfg = make_config('fixed_hip', reward='BalancingV1', reset_randomized=True, randomize_params=True,
                                randomize_gravity=True, auto_reset=True, max_episode_steps=4)
    full = oracle.Oracle(cm.struct, cfg, total, seed=5)
    full.reset()
    acts = np.random.RandomState(0).uniform(-1, 1, (9, total, 2))
    for t in range(9):
        obs, *_ = full.step(acts[t])
    assert np.array_equal(np.concatenate([r0['state'], r1['state']]), full.state)
    assert np.array_equal(np.concatenate([r0['params'], r1['params']]), full.params)
    assert np.array_equal(np.concatenate([r0['obs'], r1['obs']]), obs)
    # all-reduced statistics are identical on both ranks and equal the global sums
    assert np.array_equal(r0['red'], r1['red'])
    assert r0['red'][0] == 9 * total and r0['red'][1] == 2 * total and r0['red'][2] == 3.0
    assert r0['red'][3] == 4.0 * 2 * total
