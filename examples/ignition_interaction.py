"""The reference's examples/ignition_interaction.py on the CUDA runtime: poke ONE monopod through the ScenarIO-style
model / joint objects — reset single joints, hold a torque, advance ONE physics iteration (1e-4 s) at a time and read the
joint positions back. In the reference `gazebo.run()` advances the DART world by one iteration; here an env step of a
runtime created with agent_rate == physics_rate is exactly one physics iteration of the fused kernel (the reference's
ZOH loop `runtimes/gazebo_runtime.py:65-97` with a single sub-step). Not a hot path: every joint access is a host round
trip through os2r_get_state / os2r_set_state."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import functools

from gym_os2r import randomizers
from gym_os2r.common import make_env_from_id

# physics_rate == agent_rate: one physics iteration per env.step (the reference script builds its simulator with
# step_size=0.0001, steps_per_run=1)
make_env = functools.partial(make_env_from_id, env_id="Monopod-balance-v1", task_mode="fixed_hip",
                             agent_rate=10000, physics_rate=10000)
env = randomizers.monopod_no_rand.MonopodEnvNoRandomizer(env=make_env)
env.reset()
monopod = env.unwrapped.task.model
gazebo = env.unwrapped.gazebo
monopod.set_joint_control_mode(1, ['hip_joint', 'knee_joint'])
monopod.get_joint('hip_joint').set_joint_max_generalized_force([10])

for episode in range(2):
    # put the robot on its side, as the reference script does
    monopod.get_joint('planarizer_pitch_joint').to_gazebo().reset_position(-0.03)
    monopod.get_joint('hip_joint').to_gazebo().reset_position(-1.57)
    monopod.get_joint('knee_joint').to_gazebo().reset_position(3.14)
    gazebo.run(paused=True)
    upper_leg = monopod.get_joint('hip_joint').to_gazebo()
    lower_leg = monopod.get_joint('knee_joint').to_gazebo()
    for it in range(int(0.02 / gazebo.step_size())):          # 20 ms of physics, one iteration per step
        # the reference holds generalized force targets [0, -1] N m; actions here are torque / max_torque (2.5 N m)
        env.step([0.0, -1.0 / 2.5])
        if it % 50 == 49:
            print(f'iteration {it + 1}: lower leg {lower_leg.joint_position()[0]:+.5f} rad, '
                  f'upper leg {upper_leg.joint_position()[0]:+.5f} rad, links in contact {monopod.links_in_contact()}')
env.close()
