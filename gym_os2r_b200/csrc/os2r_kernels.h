// os2r_kernels.h — launcher interface between the C-ABI layer (os2r_capi.cu) and the kernels.
#pragma once
#include "os2r_device.cuh"

// Launch geometry of the step kernel. The fp32 kernel is compiled for 448 resident threads per SM (65 536 envs over
// 148 SMs = 443 threads/SM: ONE balanced wave; 65536/448 -> at most 144 registers per thread), either as 7 blocks of
// 2 warps (small batches: more SMs busy) or as 2 blocks of 7 warps (large batches: the lane sort of os2r_kernels.cu
// needs a few hundred envs per block to fill whole warps with one contact class).
#ifndef OS2R_BLOCK
#define OS2R_BLOCK 64
#endif
#ifndef OS2R_BLOCK_WIDE
#define OS2R_BLOCK_WIDE 224
#endif
#define OS2R_RESIDENT_THREADS 448
#ifndef OS2R_NC
#define OS2R_NC 3
#endif

namespace os2r {

template <typename T>
int step_block_threads(int64_t n_envs, int sm_count);
template <typename T>
cudaError_t launch_step(int n_dof, int n_contacts, int block, bool lone, const ModelDev<T> &M, const TaskDev &K,
                        const StateDev<T> &S, const StepIO &io, StatsDev *stats, cudaStream_t stream);
template <typename T>
cudaError_t launch_reset(int n_dof, int n_contacts, const TaskDev &K, const StateDev<T> &S, const uint8_t *mask,
                         float *obs, cudaStream_t stream);
template <typename T>
cudaError_t launch_init(const TaskDev &K, const StateDev<T> &S, double nominal_gz, cudaStream_t stream);
template <typename T>
cudaError_t step_kernel_attributes(int n_dof, int block, bool damped, cudaFuncAttributes *attr, int *blocks_per_sm);
template <typename T>
cudaError_t prepare_step(int n_dof, int block);
cudaError_t launch_fma_peak(float *out, int blocks, int iters, cudaStream_t stream);

}  // namespace os2r
