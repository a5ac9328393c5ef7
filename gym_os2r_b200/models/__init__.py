"""Model data of the monopod: URDF assets, YAML settings, and the host model compiler that turns a
URDF into the constant tables the CUDA step path consumes (reference: gym_os2r/models/)."""
from . import assets, config
from . import compiler

__all__ = ['assets', 'config', 'compiler']
