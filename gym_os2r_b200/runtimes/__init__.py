from . import configure, cuda_runtime, engine, shims

__all__ = ['cuda_runtime', 'engine', 'configure', 'shims']
