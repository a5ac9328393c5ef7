"""Minimal stand-in for the parts of OpenAI ``gym`` the reference touches (gym is not installed in
this image): ``spaces.Box``, ``Env``, ``Wrapper``, ``register`` / ``make`` / ``registry`` and the
``TimeLimit`` bookkeeping. If a real ``gym`` is importable it is used instead (see ``get_gym``).

Reference usage: gym_os2r/__init__.py:11-128 (register), gym_os2r/common/__init__.py:12-24 (make),
gym_os2r/tasks/monopod.py:124,187,198 (spaces.Box), tests/tests_general.py:7-8.
"""
import importlib
import types

import numpy as np


class Box:
    """Axis-aligned box space; ``contains`` follows gym.spaces.Box (dtype-castable, shape, bounds)."""

    def __init__(self, low, high, shape=None, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        low = np.asarray(low, dtype=np.float64)
        high = np.asarray(high, dtype=np.float64)
        if shape is not None:
            low = np.broadcast_to(low, shape)
            high = np.broadcast_to(high, shape)
        self.low = low.astype(self.dtype).copy()
        self.high = high.astype(self.dtype).copy()
        self.shape = self.low.shape
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1e6)
        hi = np.where(np.isfinite(self.high), self.high, 1e6)
        return self.np_random.uniform(lo, hi).astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        if not np.can_cast(x.dtype, self.dtype) and x.dtype.kind not in 'fiu':
            return False
        if x.shape != self.shape:
            return False
        return bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __contains__(self, x):
        return self.contains(x)

    def __repr__(self):
        return f'Box({self.low}, {self.high}, {self.shape}, {self.dtype})'

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and \
            np.allclose(self.low, other.low) and np.allclose(self.high, other.high)


class Env:
    metadata = {'render.modes': []}
    reward_range = (-float('inf'), float('inf'))
    spec = None
    action_space = None
    observation_space = None

    @property
    def unwrapped(self):
        return self

    def seed(self, seed=None):
        return [seed]

    def render(self, mode='human'):
        return None

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith('_') or name == 'env':
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    @property
    def action_space(self):
        return self.env.action_space

    @property
    def observation_space(self):
        return self.env.observation_space

    @property
    def metadata(self):
        return self.env.metadata

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def seed(self, seed=None):
        return self.env.seed(seed)

    def render(self, mode='human', **kwargs):
        return self.env.render(mode, **kwargs)

    def close(self):
        return self.env.close()


class EnvSpec:
    def __init__(self, id, entry_point, max_episode_steps=None, kwargs=None):
        self.id = id
        self.entry_point = entry_point
        self.max_episode_steps = max_episode_steps
        self.kwargs = dict(kwargs or {})

    def make(self, **overrides):
        kw = dict(self.kwargs)
        kw.update(overrides)
        ep = self.entry_point
        if isinstance(ep, str):
            mod, attr = ep.split(':')
            ep = getattr(importlib.import_module(mod), attr)
        # The CUDA runtime applies the TimeLimit on the device; hand it the limit instead of
        # wrapping (gym.make would wrap in gym.wrappers.TimeLimit).
        if self.max_episode_steps is not None and 'max_episode_steps' not in kw:
            kw['max_episode_steps'] = self.max_episode_steps
        env = ep(**kw)
        env.spec = self
        return env


class _Registry:
    def __init__(self):
        self.env_specs = {}

    def register(self, id, **kw):
        self.env_specs[id] = EnvSpec(id, **kw)

    def all(self):
        return list(self.env_specs.values())

    def make(self, id, **kw):
        if id not in self.env_specs:
            raise KeyError(f'No registered env with id: {id}')
        return self.env_specs[id].make(**kw)


registry = _Registry()


def register(id, entry_point, max_episode_steps=None, kwargs=None):
    registry.register(id, entry_point=entry_point, max_episode_steps=max_episode_steps, kwargs=kwargs)


def make(id, **kwargs):
    return registry.make(id, **kwargs)


spaces = types.SimpleNamespace(Box=Box)
