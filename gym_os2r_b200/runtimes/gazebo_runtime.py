"""Name kept for scripts / registrations that point at ``gym_os2r.runtimes.gazebo_runtime:GazeboRuntime``
(reference entry point, gym_os2r/__init__.py:18): here it IS the CUDA runtime."""
from .cuda_runtime import CudaRuntime

GazeboRuntime = CudaRuntime
