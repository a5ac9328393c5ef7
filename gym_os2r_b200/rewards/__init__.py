"""Reward classes of the monopod tasks (interface of gym_os2r/rewards/__init__.py:9-207).

Each built-in class is a closed-form function of (obs, actions[0], actions[1]). Inside the CUDA
runtime they are evaluated by the fused step kernel (``device_reward_id`` selects the formula);
``calculate_reward`` here is the host/torch mirror used by ``get_state_info`` and by tests.
User-defined subclasses (``device_reward_id = REWARD_CUSTOM``) are evaluated batched on the
GPU through ``calculate_reward`` with torch tensors (see runtimes/cuda_runtime.py).

``obs`` may be a single observation ``[D]`` or a batch ``[N, D]``; ``actions`` is indexable with
``actions[0]`` = current and ``actions[1]`` = previous normalised action (``[2]`` or ``[N, 2]``).
"""
from abc import abstractmethod

import numpy as np

from .. import _capi
from .rewards_utils import tolerance

__all__ = ['RewardBase', 'BalancingV1', 'BalancingV2', 'BalancingV3', 'StandingV1', 'HoppingV1',
           'StraightV1', 'tolerance']

_LEG_ON_BOOM_MODES = ['free_hip', 'fixed_hip', 'fixed_hip_torque', 'fixed_hip_simple', 'fixed']


def _prod_last(x):
    return x.prod(-1) if hasattr(x, 'prod') else np.prod(x, axis=-1)


def _col(obs, idx):
    return obs[..., idx]


class RewardBase:
    """Base class: subclasses set ``supported_task_modes`` and implement ``calculate_reward``."""

    device_reward_id = _capi.REWARD_CUSTOM

    def __init__(self, observation_index: dict, normalized: bool):
        self.observation_index = observation_index
        self.normalized = normalized
        self.supported_task_modes = []
        self._all_task_modes = ['free_hip', 'fixed_hip', 'fixed', 'simple', 'fixed_hip_torque',
                                'fixed_hip_simple']

    @abstractmethod
    def calculate_reward(self, obs, actions):
        """Reward for observation(s) ``obs`` given the action history ``actions``."""

    def is_task_supported(self, task_mode: str) -> bool:
        return task_mode in self.supported_task_modes

    def get_supported_task_modes(self):
        return self.supported_task_modes

    # shared pieces -------------------------------------------------------------------------
    def _height_band(self):
        """Boom-pitch band that counts as 'up': [H, 4H], H = 0.11 rad (0.11/1.57 normalised)."""
        h = 0.11 / 1.57 if self.normalized else 0.11
        return (h, 4 * h)

    def _pitch(self, obs):
        return _col(obs, self.observation_index['planarizer_pitch_joint_pos'])


class BalancingV1(RewardBase):
    """1 while the boom pitch is inside the 'up' band, else 0."""
    device_reward_id = _capi.REWARD_BALANCING_V1

    def __init__(self, observation_index: dict, normalized: bool):
        super().__init__(observation_index, normalized)
        self.supported_task_modes = list(_LEG_ON_BOOM_MODES)

    def calculate_reward(self, obs, actions):
        return tolerance(self._pitch(obs), self._height_band())


class StandingV1(BalancingV1):
    """Stand up from the ground: same indicator as BalancingV1 (rewards/__init__.py:135-149)."""


class BalancingV2(RewardBase):
    """Up-band indicator times a penalty on control magnitude (quadratic, 0.4 at |a| = 1)."""
    device_reward_id = _capi.REWARD_BALANCING_V2

    def __init__(self, observation_index: dict, normalized: bool):
        super().__init__(observation_index, normalized)
        self.supported_task_modes = list(_LEG_ON_BOOM_MODES)

    def calculate_reward(self, obs, actions):
        up = tolerance(self._pitch(obs), self._height_band())
        small_control = tolerance(actions[0], margin=1, value_at_margin=0.4, sigmoid='quadratic')
        return up * _prod_last(small_control)


class BalancingV3(RewardBase):
    """Long-tailed up-band score times a penalty on the change of control between steps."""
    device_reward_id = _capi.REWARD_BALANCING_V3

    def __init__(self, observation_index: dict, normalized: bool):
        super().__init__(observation_index, normalized)
        self.supported_task_modes = list(_LEG_ON_BOOM_MODES)

    def calculate_reward(self, obs, actions):
        up = tolerance(self._pitch(obs), self._height_band(), margin=0.01, sigmoid='long_tail')
        smooth = tolerance(actions[0] - actions[1], margin=1, value_at_margin=0.1, sigmoid='quadratic')
        return up * _prod_last(smooth)


class HoppingV1(RewardBase):
    """Up-band indicator x smooth control x forward (yaw) speed inside [0.25, 0.3] (normalised)."""
    device_reward_id = _capi.REWARD_HOPPING_V1

    def __init__(self, observation_index: dict, normalized: bool):
        super().__init__(observation_index, normalized)
        self.supported_task_modes = list(_LEG_ON_BOOM_MODES)

    def calculate_reward(self, obs, actions):
        up = tolerance(self._pitch(obs), self._height_band())
        smooth = tolerance(actions[0] - actions[1], margin=0.1, value_at_margin=0, sigmoid='quadratic')
        speed = _col(obs, self.observation_index['planarizer_yaw_joint_vel'])
        move = tolerance(speed, bounds=(0.25, 0.3), margin=0.15, value_at_margin=0.1, sigmoid='tanh_squared')
        return up * _prod_last(smooth) * move


class StraightV1(RewardBase):
    """`simple` mode: keep hip and knee at 0 (linear falloff) with a mild control penalty."""
    device_reward_id = _capi.REWARD_STRAIGHT_V1

    def __init__(self, observation_index: dict, normalized: bool):
        super().__init__(observation_index, normalized)
        self.supported_task_modes = ['simple']

    def calculate_reward(self, obs, actions):
        a = actions[0]
        a = a if hasattr(a, 'mean') else np.asarray(a, dtype=np.float64)
        control = tolerance(a / 20, margin=1, value_at_margin=0, sigmoid='quadratic').mean(-1)
        control = (4 + control) / 5
        hip = tolerance(_col(obs, self.observation_index['hip_joint_pos']), bounds=(0, 0), margin=1, sigmoid='linear')
        knee = tolerance(_col(obs, self.observation_index['knee_joint_pos']), bounds=(0, 0), margin=1, sigmoid='linear')
        return hip * knee * control
