"""The reference's examples/multiprocessing_epochs.py: there NUM_ENVS = cpu_count() Gazebo processes, here
65 536 monopods stepped by one fused kernel per step behind the same VecEnv calls."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import time

import numpy as np

from gym_os2r import randomizers
from gym_os2r.common import make_mp_envs

NUM_ENVS = 65536
envs = make_mp_envs("Monopod-balance-v1", NUM_ENVS, 42, randomizers.monopod.MonopodEnvRandomizer, task_mode='fixed_hip')
envs.reset()
returns = np.zeros(NUM_ENVS)
episodes, steps, beg = 0, 0, time.time()
rng = np.random.RandomState(0)
while episodes < 1000 and steps < 2000:
    actions = rng.uniform(-1, 1, (NUM_ENVS, 2)).astype(np.float32)
    obs, rew, done, infos = envs.step(actions)
    returns += rew
    if done.any():
        episodes += int(done.sum())
        returns[done] = 0
    steps += 1
dt = time.time() - beg
print(f'{episodes} episodes, {steps} vector steps, {steps * NUM_ENVS / dt / 1e6:.1f} M env-steps/s (numpy in/out)')
envs.close()
