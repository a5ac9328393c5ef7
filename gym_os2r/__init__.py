"""Drop-in alias: ``import gym_os2r`` resolves to the B200-native implementation ``gym_os2r_b200`` so
that scripts written against the reference keep their imports (``from gym_os2r import randomizers``,
``from gym_os2r.common import make_mp_envs``, ``from gym_os2r.rewards import BalancingV3`` ...)."""
import importlib
import sys

import gym_os2r_b200 as _impl

for _name in ('tasks', 'models', 'randomizers', 'common', 'utils', 'runtimes', 'rewards'):
    globals()[_name] = importlib.import_module('gym_os2r_b200.' + _name)
for _full, _mod in list(sys.modules.items()):
    if _full.startswith('gym_os2r_b200.'):
        sys.modules['gym_os2r.' + _full[len('gym_os2r_b200.'):]] = _mod

make, register = _impl.make, _impl.register
__all__ = ['tasks', 'models', 'randomizers', 'common', 'utils', 'runtimes', 'rewards', 'make', 'register']
