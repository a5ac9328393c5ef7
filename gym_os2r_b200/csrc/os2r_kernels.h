// os2r_kernels.h — launcher interface between the C-ABI layer (os2r_capi.cu) and the kernels.
#pragma once
#include "os2r_device.cuh"

// Launch geometry of the step kernel. 64-thread blocks: 65 536 envs -> 1024 blocks over 148 SMs =
// 6.92 blocks/SM, so with >= 7 resident blocks/SM the whole batch is ONE balanced wave
// (7 x 64 = 448 threads/SM => at most 65536/448 = 146 -> 144 registers per thread).
#ifndef OS2R_BLOCK
#define OS2R_BLOCK 64
#endif
#ifndef OS2R_MIN_BLOCKS
#define OS2R_MIN_BLOCKS 7
#endif
#define OS2R_NC 3

namespace os2r {

template <typename T>
cudaError_t launch_step(int n_dof, int n_contacts, const ModelDev<T> &M, const TaskDev &K, const StateDev<T> &S,
                        const float *actions, float *obs, float *reward, uint8_t *done, float *term_obs,
                        int32_t *info, StatsDev *stats, cudaStream_t stream);
template <typename T>
cudaError_t launch_reset(int n_dof, int n_contacts, const TaskDev &K, const StateDev<T> &S, const uint8_t *mask,
                         float *obs, cudaStream_t stream);
template <typename T>
cudaError_t launch_init(const TaskDev &K, const StateDev<T> &S, double nominal_gz, cudaStream_t stream);
template <typename T>
cudaError_t step_kernel_attributes(int n_dof, cudaFuncAttributes *attr, int *blocks_per_sm);
cudaError_t launch_fma_peak(float *out, int blocks, int iters, cudaStream_t stream);

}  // namespace os2r
