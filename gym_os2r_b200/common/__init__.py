"""Env factories — interface of gym_os2r/common/__init__.py:12-53 (``make_env_from_id`` /
``make_mp_envs``). ``make_mp_envs`` returns ONE batched CUDA env behind the VecEnv API instead of
``nenvs`` OS processes (the process fan-out of SubprocVecEnv is exactly what the GPU batch replaces)."""
import functools

from .. import _gymshim
from .vec_env import CudaVecEnv

__all__ = ['make_env_from_id', 'make_mp_envs', 'CudaVecEnv']


def make_env_from_id(env_id: str, **kwargs):
    import gym_os2r_b200  # noqa: F401  (registers the Monopod-* ids)
    return _gymshim.make(env_id, **kwargs)


def make_mp_envs(env_id, nenvs, seed, randomizer, start_idx=0, **kwargs):
    """``nenvs`` monopods with per-env RNG streams keyed by ``start_idx + i`` (the reference seeds
    worker i with ``seed + rank``, :45-53), auto-reset on done, ``info['terminal_observation']``."""
    make_env = functools.partial(make_env_from_id, env_id=env_id, num_envs=int(nenvs), seed=int(seed),
                                 first_env_id=int(start_idx), auto_reset=True, **kwargs)
    env = randomizer(env=make_env)
    env.seed(seed)
    return CudaVecEnv(env)
