"""BASELINE config 1 — the reference's examples/minimal_example.py on the CUDA runtime: ONE monopod
(`Monopod-balance-v3`: fixed_hip_simple, BalancingV2, five reset poses, TimeLimit 10 000) under the env randomizer,
seed 42, 400-step episodes with the constant action the reference script uses (or uniform random actions with
--random), optionally with a user-defined reward class (the reference advertises this extension point in the
commented block at minimal_example.py:15-29). Nothing but the import root differs from the reference script."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import functools
import time

from gym_os2r import randomizers
from gym_os2r.common import make_env_from_id
from gym_os2r.rewards import RewardBase


class ExampleV0(RewardBase):
    """A user-defined reward: 1 per step (evaluated by the runtime on the device for batched envs)."""

    def __init__(self, observation_index: dict, normalized: bool):
        super().__init__(observation_index, normalized)
        self.supported_task_modes = ['free_hip', 'fixed_hip', 'fixed_hip_torque', 'fixed_hip_simple', 'fixed']

    def calculate_reward(self, obs, actions):
        return 1


kwargs = {'reward_class': ExampleV0} if '--custom-reward' in sys.argv else {}
env_id = "Monopod-balance-v3"
make_env = functools.partial(make_env_from_id, env_id=env_id, **kwargs)
env = randomizers.monopod.MonopodEnvRandomizer(env=make_env)
env.render('human')
env.seed(42)

epochs = 5
beg_time = time.time()
steps = 0
for epoch in range(epochs):
    observation = env.reset()
    done, total_reward, c = False, 0.0, 0
    while not done:
        action = env.action_space.sample() if '--random' in sys.argv else [-1, -1]
        observation, reward, done, info = env.step(action)
        total_reward += reward
        c += 1
        done = done or c == 400
    steps += c
    print(f"Reward episode #{epoch}: {total_reward} ({c} steps, reset pose {info['reset_orientation']})")
dt = time.time() - beg_time
print(f'{steps} env steps in {dt:.2f} s incl. start-up')
env.close()
