"""Monopod task with normalised observations — vectorised host mirror of
gym_os2r/tasks/monopod.py:15-374 (``MonopodTask``).

The Task keeps the reference's method names and meaning (``create_spaces`` / ``set_action`` /
``get_observation`` / ``get_reward`` / ``is_done`` / ``reset_task`` / ``get_info`` /
``calculate_reward`` / ``get_state_info``) but operates on N environments at once: the physics,
observation, reward and termination of a step are produced by ONE fused CUDA launch owned by the
runtime (``runtimes/cuda_runtime.py``), and the ``get_*`` methods are thin views over the
tensors that launch wrote (the reference recomputes the observation three times per step,
gazebo_runtime.py:80, monopod.py:284,328).

This module also turns the YAML space definition into the C struct ``os2r_task_cfg``
(``build_task_cfg``), including raw-unit termination thresholds that are bit-equivalent to the
reference's ``not reset_space.contains(normalised_obs)`` test.
"""
import math
import struct
import warnings
from collections import deque
from typing import Deque, Dict, Tuple

import numpy as np

from .. import _capi
from .._gymshim import spaces
from ..models.config import SettingsConfig

_EPS = float(np.finfo(float).eps)


def _next_up(x: float) -> float:
    return math.nextafter(x, math.inf)


def _bits(x: float) -> int:
    """Order-preserving integer key of a double (for bisection over representable values)."""
    (u,) = struct.unpack('<q', struct.pack('<d', x))
    return u if u >= 0 else -(u & 0x7FFFFFFFFFFFFFFF)


def _from_bits(k: int) -> float:
    u = k if k >= 0 else ((-k) | (1 << 63)) - (1 << 64)
    return struct.unpack('<d', struct.pack('<q', u))[0]


def _last_true(pred, lo: float, hi: float) -> float:
    """Largest double x in [lo, hi] with pred(x) true, pred monotone (true ... true false ... false)."""
    if not pred(lo):
        return -math.inf
    if pred(hi):
        return hi
    a, b = _bits(lo), _bits(hi)
    while b - a > 1:
        mid = (a + b) // 2
        if pred(_from_bits(mid)):
            a = mid
        else:
            b = mid
    return _from_bits(a)


class MonopodTask:
    """Vectorised monopod task. Required kwargs: ``task_mode``, ``reward_class``, ``reset_positions``
    (optional ``config``: a SettingsConfig). Extra kwargs override attributes, as in the reference.
    """

    normalized = True
    supported_task_modes = ['free_hip', 'fixed_hip', 'fixed', 'fixed_hip_torque', 'simple',
                            'fixed_hip_simple']

    def __init__(self, agent_rate: float, **kwargs):
        required = ['task_mode', 'reward_class', 'reset_positions']
        for key in required:
            if key not in kwargs:
                raise RuntimeError(f'Missing required kwarg: {key}. We require the following kwargs, '
                                   f'{required}\n in the MonopodTask class. (These can be specified in env init)')
        if len(kwargs) != len(required):
            warnings.warn(f'# WARNING: Supplied Kwargs, {kwargs} Contains more entries than expected. '
                          f'Required Kwargs are {required}. Could be caused by config object.',
                          SyntaxWarning, stacklevel=2)
        self.agent_rate = agent_rate
        self.__dict__.update(kwargs)
        self.cfg = kwargs.get('config') or SettingsConfig()
        known_resets = list(self.cfg.get_config('/resets').keys())
        if not set(self.reset_positions).issubset(known_resets):
            raise RuntimeError('One or more of the reset positions provided were not in the supported '
                               f'reset positions. {known_resets}')
        if self.task_mode not in self.supported_task_modes:
            raise RuntimeError(f'task mode {self.task_mode} not supported in monopod environment.')
        try:
            self.spaces_definition = self.cfg.get_config(f'task_modes/{self.task_mode}/spaces')
        except KeyError:
            raise RuntimeError(f'task mode {self.task_mode} does not contain spaces definition in monopod '
                               'environment config file.')

        self.model_name = None
        self.model = None          # ScenarIO-style shim set by the runtime
        self.world = None
        self.runtime = None        # CudaRuntime that owns the device state
        self.np_random = np.random.RandomState()
        self.action_space = None
        self.observation_space = None
        self.reset_space = None
        self.current_reset_orientation = None

        self.action_names = [*self.spaces_definition['action']]
        self.joint_names = [*self.spaces_definition['observation']]
        self.observation_index: Dict[str, int] = {}
        # Reference keeps a 10-deep deque of which only entries 0 and 1 are ever read; the device
        # keeps exactly those two (a_t, a_{t-1}); this host deque mirrors them for API parity.
        self.action_history: Deque = deque([np.zeros(len(self.action_names)) for _ in range(10)], maxlen=10)
        self.observing_measured_torque = self.spaces_definition['observing_measured_torque']
        self.observation_name_mask = self.spaces_definition['observation_mask']
        self.__dict__.update(kwargs)

    # ------------------------------------------------------------------ spaces
    def seed_task(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def create_spaces(self) -> Tuple[spaces.Box, spaces.Box]:
        """Build action / observation / reset spaces from the YAML definition (monopod.py:105-200)."""
        self.max_torques = np.array(list(self.spaces_definition['action'].values()), dtype=np.float64)
        action_space = spaces.Box(low=np.array([-1.0, -1.0]), high=np.array([1.0, 1.0]), dtype=np.float64)

        limits = np.array([info['limits'] for info in self.spaces_definition['observation'].values()],
                          dtype=np.float64)
        low = np.concatenate((limits[:, 1], limits[:, 3]))
        high = np.concatenate((limits[:, 0], limits[:, 2]))
        names = [n + '_pos' for n in self.joint_names] + [n + '_vel' for n in self.joint_names]
        if self.observing_measured_torque:
            low = np.concatenate((low, action_space.low))
            high = np.concatenate((high, action_space.high))
            names += [n + '_torque' for n in self.action_names]
        self.observation_names_unmasked = names

        self.observation_mask, self.velocities_index, self.observation_index = [], [], {}
        for i, name in enumerate(names):
            if name in self.observation_name_mask:
                continue
            col = len(self.observation_mask)
            self.observation_index[name] = col
            self.observation_mask.append(i)
            if '_vel' in name:
                self.velocities_index.append(col)
        low, high = low[self.observation_mask], high[self.observation_mask]

        self.periodic_joints = [self.observation_index[j + '_pos']
                                for j, info in self.spaces_definition['observation'].items()
                                if info['periodic_pos'] and j + '_pos' in self.observation_index]
        low[self.periodic_joints] = -(np.pi + _EPS)
        high[self.periodic_joints] = np.pi + _EPS
        self.obs_limits = {'high': high.copy(), 'low': low.copy()}
        self.mask_inf_obs = np.zeros(len(high), dtype=bool)
        self.mask_inf_obs[self.velocities_index] = True
        if self.normalized:
            low = np.full_like(low, -1.0)
            high = np.full_like(high, 1.0)
        obs_space = spaces.Box(low=low, high=high, dtype=np.float64)

        self.reward = self.reward_class(self.observation_index, normalized=self.normalized)
        assert self.reward.is_task_supported(self.task_mode), \
            f"'{self.task_mode}' task mode not supported by reward class '{self.reward}'"
        self.reset_space = spaces.Box(low=low + _EPS, high=high - _EPS, dtype=np.float64)
        self.action_space, self.observation_space = action_space, obs_space
        return action_space, obs_space

    # ------------------------------------------------------------------ host formulas
    def observation_from_raw(self, joint_pos, joint_vel, prev_action=None):
        """Reference observation formula on host arrays ``[..., n_joints]`` (joint_names order);
        monopod.py:238-272. Used by tests / shims; the step path computes this on the device."""
        parts = [np.asarray(joint_pos, dtype=np.float64), np.asarray(joint_vel, dtype=np.float64)]
        if self.observing_measured_torque:
            parts.append(np.asarray(prev_action, dtype=np.float64))
        obs = np.concatenate(parts, axis=-1)[..., self.observation_mask]
        pj = self.periodic_joints
        obs[..., pj] = np.mod(obs[..., pj] + np.pi, 2 * np.pi) - np.pi
        if self.normalized:
            high, low, m = self.obs_limits['high'], self.obs_limits['low'], self.mask_inf_obs
            obs[..., ~m] = 2 * (obs[..., ~m] - low[~m]) / (high[~m] - low[~m]) - 1
            obs[..., m] = np.tanh(0.05 * obs[..., m])
        return obs

    @staticmethod
    def _history(actions):
        """Normalise ``actions`` to ``[a_t, a_{t-1}]``. The reference documents a deque of past
        actions but its callers pass a bare action (examples/fixed.py:45, tests/tests_general.py:92);
        both are accepted: a deque / list of arrays is a history, anything else is a bare action
        (or an ``[N, 2]`` batch of them) and stands for the history ``[a, a]``."""
        if isinstance(actions, deque):
            return [actions[0], actions[1]]
        if isinstance(actions, (list, tuple)) and len(actions) >= 2 and np.ndim(actions[0]) >= 1:
            return [actions[0], actions[1]]
        a = actions if hasattr(actions, 'dim') else np.asarray(actions, dtype=np.float64)
        return [a, a]

    def calculate_reward(self, obs, actions):
        return self.reward.calculate_reward(obs, actions)

    def done_from_observation(self, obs):
        """``not reset_space.contains(obs)`` per row (monopod.py:284-286)."""
        if hasattr(obs, 'dim'):
            import torch
            lo = torch.as_tensor(self.reset_space.low, dtype=obs.dtype, device=obs.device)
            hi = torch.as_tensor(self.reset_space.high, dtype=obs.dtype, device=obs.device)
            return ~(((obs >= lo) & (obs <= hi)).all(-1))
        o = np.asarray(obs, dtype=np.float64)
        return ~np.logical_and(o >= self.reset_space.low, o <= self.reset_space.high).all(-1)

    def get_state_info(self, obs, actions):
        """Stateless (reward, done) of an observation — fixes the reference's NameError
        (monopod.py:364 uses the undefined name ``action``)."""
        reward = self.calculate_reward(obs, self._history(actions))
        done = self.done_from_observation(obs)
        if np.ndim(done) == 0:
            return float(reward), bool(done)
        return reward, done

    # ------------------------------------------------------------------ device-backed Task API
    def set_action(self, action, store_action: bool = True) -> bool:
        """Stage the torque command for the next fused step (monopod.py:202-236). The zero-order
        hold over the 10 physics iterations and the action-history update happen in the kernel."""
        if self.runtime is None:
            raise RuntimeError('task is not attached to a runtime')
        self.runtime._stage_action(action)
        return True

    def get_observation(self):
        return self.runtime._last('obs')

    def get_reward(self):
        return self.runtime._last('reward')

    def is_done(self):
        return self.runtime._last('done')

    def reset_task(self) -> None:
        """Force-control mode and max torque are constants of the CUDA backend (monopod.py:300-318)."""
        if self.runtime is None:
            raise RuntimeError('task is not attached to a runtime')

    def get_info(self) -> Dict:
        return {'reset_orientation': self.current_reset_orientation}


# ---------------------------------------------------------------------------------------------
# YAML definition -> struct os2r_task_cfg
# ---------------------------------------------------------------------------------------------

RANDOMIZATION_DEFAULTS = dict(   # gym_os2r/randomizers/monopod.py:182-215,58
    mass_lo=0.8, mass_hi=1.2, fric_lo=0.01, fric_hi=0.05, damp_lo=0.8, damp_hi=1.2,
    mu_lo=0.8, mu_hi=1.2, mu_link=0.33, grav_mean=-9.8, grav_std=0.2)


def build_task_cfg(task: MonopodTask, compiled_model, *, max_episode_steps: int, auto_reset: bool,
                   reset_randomized: bool, randomize_params: bool, randomize_gravity: bool,
                   randomization: dict = None, gravity_redraw_resets: int = 0) -> _capi.TaskCfg:
    """Translate a created task (``create_spaces`` already called) into the device configuration."""
    t = _capi.TaskCfg()
    D = len(task.observation_mask)
    if D > _capi.MAX_OBS:
        raise ValueError('observation too wide')
    t.obs_dim = D
    t.normalized = int(task.normalized)
    t.reward_id = int(getattr(task.reward, 'device_reward_id', _capi.REWARD_CUSTOM))
    t.max_episode_steps = int(max_episode_steps or 0)
    t.auto_reset = int(auto_reset)
    t.reset_randomized = int(reset_randomized)
    t.randomize_params = int(randomize_params)
    t.randomize_gravity = int(randomize_gravity)
    t.simple_sample_reset = int(task.task_mode == 'simple' and not reset_randomized)
    if int(gravity_redraw_resets) < 0:
        raise ValueError('num_physics_rollouts must be >= 0')
    t.gravity_redraw_resets = int(gravity_redraw_resets)   # MonopodEnvRandomizer(num_physics_rollouts=K)

    nj = len(task.joint_names)
    col = lambda name: task.observation_index.get(name, -1)
    t.reward_pitch_col = col('planarizer_pitch_joint_pos')
    t.reward_yawvel_col = col('planarizer_yaw_joint_vel')
    t.reward_hip_col = col('hip_joint_pos')
    t.reward_knee_col = col('knee_joint_pos')
    needs = {_capi.REWARD_BALANCING_V1: ['reward_pitch_col'], _capi.REWARD_BALANCING_V2: ['reward_pitch_col'],
             _capi.REWARD_BALANCING_V3: ['reward_pitch_col'],
             _capi.REWARD_HOPPING_V1: ['reward_pitch_col', 'reward_yawvel_col'],
             _capi.REWARD_STRAIGHT_V1: ['reward_hip_col', 'reward_knee_col']}
    for f in needs.get(t.reward_id, []):
        if getattr(t, f) < 0:
            raise KeyError(f'reward {type(task.reward).__name__} needs an observation column that is masked out ({f})')

    for k, src in enumerate(task.observation_mask):
        lo, hi = float(task.obs_limits['low'][k]), float(task.obs_limits['high'][k])
        t.obs_low[k], t.obs_high[k] = lo, hi
        if src < nj:
            jname = task.joint_names[src]
            periodic = k in task.periodic_joints
            t.obs_kind[k] = _capi.OBS_POS_PERIODIC if periodic else _capi.OBS_POS
            t.obs_index[k] = compiled_model.dof_of(jname)
        elif src < 2 * nj:
            t.obs_kind[k] = _capi.OBS_VEL
            t.obs_index[k] = compiled_model.dof_of(task.joint_names[src - nj])
        else:
            t.obs_kind[k] = _capi.OBS_TORQUE
            t.obs_index[k] = src - 2 * nj
        t.done_low[k], t.done_high[k] = _raw_done_thresholds(task, k, t.obs_kind[k], lo, hi)

    resets = task.cfg.get_config('/resets')
    if len(task.reset_positions) > _capi.MAX_RESETS:
        raise ValueError('too many reset positions')
    t.n_resets = len(task.reset_positions)
    for i, name in enumerate(task.reset_positions):
        t.reset_pitch[i] = float(resets[name]['planarizer_pitch_joint'])
        t.reset_laying[i] = int(bool(resets[name]['laying_down']))
    d = task.cfg.get_config(f'task_modes/{task.task_mode}/definition')
    t.ik_upper_leg, t.ik_lower_leg = float(d['upper_leg_length']), float(d['lower_leg_length'])
    t.ik_pivot_height, t.ik_boom = float(d['central_pivot_height']), float(d['length_boom'])
    t.ik_hip_offset, t.ik_clip = float(d['hip_offset']), float(d['clipping_adjust'])
    if t.simple_sample_reset:
        for j, name in enumerate(('hip_joint_pos', 'knee_joint_pos')):
            c = task.observation_index[name]
            t.simple_lo[j] = float(task.observation_space.low[c])
            t.simple_hi[j] = float(task.observation_space.high[c])
    r = dict(RANDOMIZATION_DEFAULTS)
    r.update(randomization or {})
    for key, val in r.items():
        setattr(t, key, float(val))
    return t


def _raw_done_thresholds(task: MonopodTask, k: int, kind: int, lo: float, hi: float):
    """Raw-unit interval [a, b] such that  (a <= raw <= b)  <=>  reset_space.contains(column k).

    Found by bisection over representable doubles on the *exact* reference arithmetic
    (2*(x-low)/(high-low)-1 resp. tanh(0.05 x)), so that the device can decide termination on raw
    state in fp64 and agree bit-for-bit with the reference's test on the normalised observation
    (fp32 tanh would saturate at |v| ~ 173 rad/s instead of ~ 367).
    """
    rs_lo, rs_hi = float(task.reset_space.low[k]), float(task.reset_space.high[k])
    if kind == _capi.OBS_POS_PERIODIC:
        # The wrapped value lies in [-pi, pi]; the column only terminates within a few ulp of the
        # wrap point (normalised -1 < -1+eps). The device wraps in fp64 (fmod is exact), so even
        # this corner agrees with the reference.
        if task.normalized:
            f = lambda x: 2 * (x - lo) / (hi - lo) - 1
            a = -_last_true(lambda y: f(-y) >= rs_lo, 0.0, math.pi)
            b = _last_true(lambda x: f(x) <= rs_hi, 0.0, math.pi)
            return a, b
        return rs_lo, rs_hi
    if not task.normalized:
        return rs_lo, rs_hi          # raw comparison already (±inf for velocities)
    if kind == _capi.OBS_VEL:
        f = lambda x: math.tanh(0.05 * x)
        b = _last_true(lambda x: f(x) <= rs_hi, 0.0, 1e6)
        a = -_last_true(lambda y: f(-y) >= rs_lo, 0.0, 1e6)
        return a, b
    f = lambda x: 2 * (x - lo) / (hi - lo) - 1
    mid = 0.5 * (lo + hi)
    span = 2.0 * (hi - lo)
    b = _last_true(lambda x: f(x) <= rs_hi, mid, mid + span)
    a = -_last_true(lambda y: f(-y) >= rs_lo, -mid, -mid + span)
    return a, b
