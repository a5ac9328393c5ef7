"""gym_os2r_b200 — B200-native simulation backend for the OpenSim2Real monopod environments.

Same task ids, Task / rewards / randomizers interfaces as gym-os2r (reference: gym_os2r/__init__.py:16-128),
with the Ignition-Gazebo/DART runtime replaced by ``runtimes.cuda_runtime.CudaRuntime``.
"""
from . import _gymshim
from . import common, models, randomizers, rewards, runtimes, tasks, utils
from ._gymshim import make, register
from .rewards import BalancingV1, BalancingV2, BalancingV3, HoppingV1, StandingV1, StraightV1  # noqa: F401

__all__ = ['tasks', 'models', 'randomizers', 'common', 'utils', 'runtimes', 'rewards', 'make', 'register']

ENTRY_POINT = 'gym_os2r_b200.runtimes.cuda_runtime:CudaRuntime'
_COMMON = dict(agent_rate=1000, physics_rate=10000, real_time_factor=3.4028234663852886e+38)


def _reg(env_id, task_cls, task_mode, reward_class, reset_positions, max_episode_steps=100_000):
    if env_id in _gymshim.registry.env_specs:
        return
    register(id=env_id, entry_point=ENTRY_POINT, max_episode_steps=max_episode_steps,
             kwargs=dict(task_cls=task_cls, task_mode=task_mode, reward_class=reward_class,
                         reset_positions=list(reset_positions), **_COMMON))


_ALL_POSES = ['stand', 'half_stand', 'ground', 'lay', 'float']
_reg('Monopod-stand-v1', tasks.monopod.MonopodTask, 'fixed_hip', StandingV1, ['ground'])
_reg('Monopod-balance-v1', tasks.monopod.MonopodTask, 'fixed_hip_simple', BalancingV1, ['stand'])
_reg('Monopod-balance-v2', tasks.monopod.MonopodTask, 'fixed_hip_simple', BalancingV2, ['stand'])
_reg('Monopod-balance-v3', tasks.monopod.MonopodTask, 'fixed_hip_simple', BalancingV2, _ALL_POSES, 10_000)
_reg('Monopod-nonorm-balance-v1', tasks.monopod_no_norm.MonopodTask, 'fixed_hip_simple', BalancingV1, ['stand'])
_reg('Monopod-nonorm-balance-v2', tasks.monopod_no_norm.MonopodTask, 'fixed_hip_simple', BalancingV2, ['stand'])
_reg('Monopod-nonorm-balance-v3', tasks.monopod_no_norm.MonopodTask, 'fixed_hip_simple', BalancingV2, _ALL_POSES, 10_000)
_reg('Monopod-hop-v1', tasks.monopod.MonopodTask, 'free_hip', HoppingV1, ['stand'])
_reg('Monopod-simple-v1', tasks.monopod.MonopodTask, 'simple', StraightV1, ['stand'])
