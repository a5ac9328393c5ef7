"""Single monopod, fixed boom (hanging leg, 3 DoF), random actions — the reference's examples/fixed.py (BASELINE config 2) on the CUDA runtime."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import functools
import time

from gym_os2r import randomizers
from gym_os2r.common import make_env_from_id

env_id = "Monopod-balance-v1"
make_env = functools.partial(make_env_from_id, env_id=env_id, task_mode='fixed')
env = randomizers.monopod.MonopodEnvRandomizer(env=make_env)
env.render('human')
env.seed(42)

beg_time = time.time()
for epoch in range(3):
    observation = env.reset()
    done, total_reward, steps = False, 0.0, 0
    while not done and steps < 2000:
        action = env.action_space.sample()
        observation, reward, done, _ = env.step(action)
        total_reward += reward
        steps += 1
    print(f"Reward episode #{epoch}: {total_reward} ({steps} steps); rollout info:",
          env.get_state_info(observation, action))
print('time:', time.time() - beg_time)
env.close()
