from .cuda_vec_env import CudaVecEnv

__all__ = ['CudaVecEnv']
