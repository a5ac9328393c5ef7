from . import configure, cuda_runtime, engine, gazebo_runtime, shims

__all__ = ['cuda_runtime', 'gazebo_runtime', 'engine', 'configure', 'shims']
