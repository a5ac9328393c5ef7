"""``CudaRuntime``: the gym.Env runtime that replaces the reference's ``GazeboRuntime``
(gym_os2r/runtimes/gazebo_runtime.py:12-123) by one fused CUDA launch per env step.

Same constructor signature as the reference (``task_cls, agent_rate, physics_rate,
real_time_factor, physics_engine, world, **task kwargs``) plus ``num_envs`` / ``device`` / ``seed``.
Switching an existing script means changing the registered ``entry_point`` (done in
``gym_os2r_b200/__init__.py``) and nothing else.

Shapes
  * ``num_envs == 1`` (default): the reference's single-env API — ``step(action[2])`` returns
    ``(obs float64[D], float reward, bool done, dict info)``; ``reset()`` returns ``obs``.
  * ``num_envs > 1``: batched — CUDA tensors ``obs[N, D]`` float32, ``reward[N]``, ``done[N]`` bool and an
    ``info`` dict of tensors (``reset_orientation`` ids, ``terminal_observation``, ``cause`` bits).

The physics engine handle is created lazily (first ``reset``/``step``/``seed``) so that the env
randomizer wrappers, which in the reference own the reset logic, can first configure how resets and
parameter draws behave (``configure_randomization``).
"""
import warnings
from typing import Optional

import numpy as np
import torch

from .. import _capi
from .._gymshim import Env
from .configure import configure
from .engine import Engine
from .shims import ModelShim, SimulatorShim, WorldShim


class CudaRuntime(Env):
    metadata = {'render.modes': ['human']}

    def __init__(self, task_cls: type, agent_rate: float, physics_rate: float, real_time_factor: float = None,
                 physics_engine=None, world: str = None, num_envs: int = 1, device: int = 0,
                 seed: Optional[int] = None, max_episode_steps: int = 0, auto_reset: Optional[bool] = None,
                 precision: int = 32, first_env_id: int = 0, pgs_iters: Optional[int] = None, pgs_tol: Optional[float] = None,
                 tuning: Optional[dict] = None, **kwargs):
        steps = physics_rate / agent_rate
        if steps != int(steps):
            warnings.warn(f'Rounding the number of iterations to {int(steps)} from the nominal {steps}')
        self.num_of_steps_per_run = int(steps)
        self.agent_rate, self.physics_rate, self.real_time_factor = agent_rate, physics_rate, real_time_factor
        self.num_envs = int(num_envs)
        self.batched = self.num_envs > 1
        self.device_index = int(device)
        self._seed = 0 if seed is None else int(seed)
        self._precision = int(precision)
        self._first_env_id = int(first_env_id)
        self._tuning = dict(tuning) if tuning else None
        self._task_cls, self._task_kwargs = task_cls, dict(kwargs)
        self._opts = dict(max_episode_steps=int(max_episode_steps or 0),
                          auto_reset=self.batched if auto_reset is None else bool(auto_reset),
                          reset_randomized=False, randomize_params=False, randomize_gravity=False,
                          randomization=None, gravity_redraw_resets=0, pgs_iters=pgs_iters, pgs_tol=pgs_tol)
        self._engine: Optional[Engine] = None
        self._build_task()
        self._gazebo = SimulatorShim(self)
        self._world = WorldShim(self)
        self._out = {}
        self._staged = None
        self._prev_actions = None
        self._custom_returns = None       # per-episode returns of a user-defined reward (device tensor)
        self._custom_sum_return = None
        self.spec = None

    # ------------------------------------------------------------------ configuration
    def _build_task(self):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', SyntaxWarning)
            self.task, self._compiled, self._cfg = configure(
                self._task_cls, self.agent_rate, self.physics_rate, **self._opts, **self._task_kwargs)
        self.task.runtime = self
        self.task.world = getattr(self, '_world', None)
        self.task.model = ModelShim(self)
        self.task.model_name = self.task.model.name()
        self.action_space, self.observation_space = self.task.action_space, self.task.observation_space
        self.task.seed_task(self._seed)
        self.action_space.seed(self._seed)

    def configure_randomization(self, *, reset_randomized=None, randomize_params=None, randomize_gravity=None,
                                randomization=None, gravity_redraw_resets=None):
        """Called by the env randomizer wrappers before the first reset (reference: the wrappers own
        ``randomize_task`` / ``randomize_physics`` / ``randomize_model_description``). On a live engine the new
        ranges / switches are applied in place (``os2r_set_randomization``) and take effect at the next resets."""
        changed = False
        for key, val in (('reset_randomized', reset_randomized), ('randomize_params', randomize_params),
                         ('randomize_gravity', randomize_gravity), ('randomization', randomization),
                         ('gravity_redraw_resets', gravity_redraw_resets)):
            if val is not None and self._opts[key] != val:
                self._opts[key] = val
                changed = True
        if changed:
            self._build_task()
            if self._engine is not None:
                self._engine.set_randomization(self._cfg)
                self._engine.task_cfg = self._cfg

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self._compiled, self._cfg, self.num_envs, device=self.device_index,
                                  seed=self._seed, first_env_id=self._first_env_id, precision=self._precision,
                                  tuning=self._tuning)
        return self._engine

    # properties the reference exposes
    @property
    def gazebo(self):
        return self._gazebo

    @property
    def world(self):
        return self._world

    # ------------------------------------------------------------------ gym.Env
    def seed(self, seed: int = None):
        self._seed = 0 if seed is None else int(seed)
        self.task.seed_task(self._seed)
        self.action_space.seed(self._seed)
        if self._engine is not None:
            self._engine.seed(self._seed)
        return [self._seed]

    def _reset_names(self, ids):
        names = self.task.reset_positions
        return [names[int(i)] for i in ids]

    def reset(self, mask=None):
        eng = self.engine
        if mask is not None and not torch.is_tensor(mask):
            mask = torch.as_tensor(np.asarray(mask), device=eng.device)
        obs = eng.reset(mask)
        if self.batched:
            return obs
        torch.cuda.synchronize(eng.device)
        rid = int(eng.get_reset_ids()[0])
        self.task.current_reset_orientation = self.task.reset_positions[rid] if rid < len(self.task.reset_positions) else None
        return obs[0].double().cpu().numpy()

    def _stage_action(self, action):
        self._staged = action

    def step(self, action):
        eng = self.engine
        custom = self._cfg.reward_id == _capi.REWARD_CUSTOM
        if not self.batched and not custom:
            return self._step_single(action)
        if self.batched:
            a = action if torch.is_tensor(action) else torch.as_tensor(np.asarray(action, dtype=np.float32), device=eng.device)
            a = a.to(device=eng.device, dtype=torch.float32).reshape(self.num_envs, 2).contiguous()
        else:
            arr = np.asarray(action, dtype=np.float64).reshape(2)
            if not self.action_space.contains(arr):
                warnings.warn('The action does not belong to the action space')
            a = torch.as_tensor(arr.astype(np.float32), device=eng.device).reshape(1, 2)
        obs, reward, done_u8, info_t = eng.step(a)
        if custom:
            # User-defined RewardBase subclass: one batched torch evaluation on the device, on the observation the
            # step produced BEFORE any auto-reset (terminal_obs equals obs for envs that did not finish). The kernel
            # accumulated reward 0, so the per-episode returns and their statistics are kept here.
            prev = self._prev_actions if self._prev_actions is not None else torch.zeros_like(a)
            r = self.task.reward.calculate_reward(eng.terminal_obs.double(), [a.double(), prev.double()])
            reward = torch.as_tensor(r, device=eng.device, dtype=torch.float32).expand(self.num_envs).contiguous()
            eng.reward.copy_(reward)
            if self._custom_returns is None:
                self._custom_returns = torch.zeros(self.num_envs, dtype=torch.float64, device=eng.device)
                self._custom_sum_return = torch.zeros((), dtype=torch.float64, device=eng.device)
            self._custom_returns += reward.double()
            fin = done_u8.bool()
            self._custom_sum_return += (self._custom_returns * fin).sum()
            self._custom_returns.masked_fill_(fin, 0.0)
        self._prev_actions = a.clone()
        self._out = {'obs': obs, 'reward': reward, 'done': done_u8.bool()}
        if self.batched:
            info = {'reset_orientation': info_t[:, 0], 'cause': info_t[:, 1],
                    'terminal_observation': eng.terminal_obs,
                    'TimeLimit.truncated': (info_t[:, 1] & 3) == 2}
            return obs, reward, self._out['done'], info
        torch.cuda.synchronize(eng.device)
        info_h = info_t[0].cpu().numpy()
        self.task.action_history.appendleft(a[0].double().cpu().numpy())
        return self._single_result(obs[0].double().cpu().numpy(), float(reward[0].item()), bool(done_u8[0].item()), info_h)

    def _step_single(self, action):
        """The reference's single-env call shape: one C-ABI host call (H2D + kernel + D2H), no torch round trips."""
        arr = np.asarray(action, dtype=np.float64).reshape(2)
        if not self.action_space.contains(arr):
            warnings.warn('The action does not belong to the action space')
        obs, rew, done, _, info = self.engine.step_host(arr.astype(np.float32).reshape(1, 2), False, True)
        self.task.action_history.appendleft(arr.copy())
        self._out = {'obs_h': obs[0].astype(np.float64), 'reward_h': float(rew[0]), 'done_h': bool(done[0])}
        return self._single_result(self._out['obs_h'], self._out['reward_h'], self._out['done_h'], info[0])

    def _single_result(self, obs, reward, done, info_row):
        self.task.current_reset_orientation = self.task.reset_positions[int(info_row[0])]
        info = self.task.get_info()
        if (int(info_row[1]) & 3) == 2:
            info['TimeLimit.truncated'] = True
        return obs, reward, done, info

    def _last(self, key):
        """Outputs of the most recent fused step (what the Task's get_* methods return)."""
        if key + '_h' in self._out:
            return self._out[key + '_h']
        if key not in self._out:
            raise RuntimeError('no step has been executed yet')
        v = self._out[key]
        if self.batched:
            return v
        if key == 'obs':
            return v[0].double().cpu().numpy()
        return float(v[0].item()) if key == 'reward' else bool(v[0].item())

    def get_state_info(self, obs, actions):
        return self.task.get_state_info(obs, actions)

    def render(self, mode: str = 'human', **kwargs):
        if mode not in self.metadata['render.modes']:
            raise ValueError(f'Unsupported render mode {mode!r}')
        return True   # headless backend: rendering is a no-op (no GUI process to attach to)

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def stats(self, clear: bool = False) -> dict:
        """Device-side episode statistics (``os2r_stats_read``). With a user-defined reward class the returns are
        accumulated by this runtime (the kernel saw reward 0) and substituted here."""
        st = self.engine.stats(clear)
        if self._custom_sum_return is not None:
            st['sum_return'] = float(self._custom_sum_return.item())
            if clear:
                self._custom_sum_return.zero_()
        return st

    # ------------------------------------------------------------------ checkpoint / resume
    def get_state(self):
        """Everything needed to continue the rollout in a NEW runtime: joint state + warm-start impulses + last
        action, the randomised parameters, and the per-env bookkeeping — TimeLimit clocks, episode returns, reset
        ids and the episode counters that key each env's RNG stream — plus the seed and the statistics."""
        eng = self.engine
        steps, ret = eng.get_episode()
        snap = {'state': eng.get_state(), 'params': eng.get_params(), 'steps': steps, 'returns': ret,
                'reset_ids': eng.get_reset_ids(), 'episodes': eng.get_episode_counters(), 'seed': self._seed,
                'stats': eng.stats()}
        if self._custom_returns is not None:
            snap['custom_returns'] = self._custom_returns.cpu().numpy()
            snap['custom_sum_return'] = float(self._custom_sum_return.item())
        if self._prev_actions is not None:
            snap['prev_actions'] = self._prev_actions.cpu().numpy()
        return snap

    def set_state(self, snapshot):
        eng = self.engine
        if 'seed' in snapshot:
            self.seed(int(snapshot['seed']))
        eng.set_state(snapshot['state'])
        if 'params' in snapshot:
            eng.set_params(snapshot['params'])
        eng.set_episode(snapshot.get('steps'), snapshot.get('returns'), snapshot.get('reset_ids'),
                        snapshot.get('episodes'))
        if 'stats' in snapshot:
            eng.set_stats(snapshot['stats'])
        if 'custom_returns' in snapshot:
            self._custom_returns = torch.as_tensor(snapshot['custom_returns'], device=eng.device, dtype=torch.float64).clone()
            self._custom_sum_return = torch.tensor(snapshot['custom_sum_return'], device=eng.device, dtype=torch.float64)
        if 'prev_actions' in snapshot:
            self._prev_actions = torch.as_tensor(snapshot['prev_actions'], device=eng.device, dtype=torch.float32).clone()
