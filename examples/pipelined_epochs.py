"""`VecEnv.step_async` / `step_wait` (reference: common/vec_env/subproc_vec_env.py:114-123) used the way the split exists
for: TWO env groups on one GPU, each behind its own VecEnv (own C-ABI handle and stream); while group A's results travel
to the host and its next actions are produced, group B's kernel runs. The numpy-facing path is PCIe-bound (3 MB per step of
65 536 envs), so overlapping the copies of one group with the kernel of the other raises the end-to-end rate."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import time

import numpy as np

from gym_os2r import randomizers
from gym_os2r.common import make_mp_envs

HALF = 32768
groups = [make_mp_envs("Monopod-balance-v1", HALF, 42, randomizers.monopod.MonopodEnvRandomizer, start_idx=k * HALF,
                       task_mode='fixed_hip') for k in range(2)]
for g in groups:
    g.reset()
rng = np.random.RandomState(0)
pool = [rng.uniform(-1, 1, (HALF, 2)).astype(np.float32) for _ in range(16)]


def policy(k, t, obs):
    """stand-in for a host-side policy: writes the next actions into the group's page-locked action buffer"""
    groups[k].action_buffer[:] = pool[(2 * t + k) % 16]
    return groups[k].action_buffer


steps = 1000
for k in range(2):                      # prime the pipeline
    groups[k].step_async(policy(k, 0, None))
t0 = time.perf_counter()
for t in range(1, steps + 1):
    for k in range(2):
        obs, rew, done, infos = groups[k].step_wait()          # group k's results (the other group's step is in flight)
        groups[k].step_async(policy(k, t, obs))
dt = time.perf_counter() - t0
for k in range(2):
    groups[k].step_wait()
print(f'pipelined: {2 * HALF * steps / dt / 1e6:.1f} M env-steps/s end to end (numpy in / out, 2 x {HALF} envs, one GPU)')

one = make_mp_envs("Monopod-balance-v1", 2 * HALF, 42, randomizers.monopod.MonopodEnvRandomizer, task_mode='fixed_hip')
one.reset()
big = [np.concatenate([pool[i], pool[(i + 1) % 16]]) for i in range(16)]
for t in range(50):
    one.action_buffer[:] = big[t % 16]
    one.step(one.action_buffer)
t0 = time.perf_counter()
for t in range(steps):
    one.action_buffer[:] = big[t % 16]
    obs, rew, done, infos = one.step(one.action_buffer)
dt = time.perf_counter() - t0
print(f'plain    : {2 * HALF * steps / dt / 1e6:.1f} M env-steps/s end to end (one VecEnv of {2 * HALF} envs)')
for g in groups + [one]:
    g.close()
