"""The reference's examples/minimal_example_testing.py on the CUDA runtime: a user-defined reset pose written into a
SettingsConfig (`resets/<name>/planarizer_pitch_joint`, `.../laying_down`) and handed to the task through the `config`
kwarg together with `reset_positions=[<name>]` (reference :17-31)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import functools

from gym_os2r import randomizers
from gym_os2r.common import make_env_from_id
from gym_os2r.models.config import SettingsConfig

reset_position = 'name_reset'
cfg = SettingsConfig()
cfg.set_config(True, 'resets/' + reset_position + '/laying_down')
cfg.set_config(0.4, 'resets/' + reset_position + '/planarizer_pitch_joint')

make_env = functools.partial(make_env_from_id, env_id="Monopod-stand-v1", reset_positions=[reset_position], config=cfg)
env = randomizers.monopod_no_rand.MonopodEnvNoRandomizer(env=make_env)
env.render('human')
env.seed(42)
for epoch in range(3):
    observation = env.reset()
    done, total, c = False, 0.0, 0
    while not done and c < 300:
        observation, reward, done, info = env.step([-1, -1])
        total += reward
        c += 1
    pitch = env.unwrapped.task.model.joint_positions(['planarizer_pitch_joint'])[0]
    print(f"episode #{epoch}: reset pose {info['reset_orientation']!r} (boom released at 0.4 rad, laying), return {total}, "
          f"{c} steps, boom pitch now {pitch:+.3f} rad")
env.close()
