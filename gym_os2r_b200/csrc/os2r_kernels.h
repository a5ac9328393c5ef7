// os2r_kernels.h — launcher interface between the C-ABI layer (os2r_capi.cu) and the kernels.
#pragma once
#include "os2r_device.cuh"

// Launch geometry of the step kernel. The fp32 product build steps one env per thread: 65 536 envs over 148 SMs = 443
// threads/SM = ONE balanced wave, either as 2 blocks of 7 warps at 128 registers per thread (large batches: the lane
// sort of os2r_kernels.cu needs a few hundred envs per block to fill whole warps with one contact class) or as 2-warp
// blocks without an occupancy target (small batches: more SMs busy, ~200 registers per thread).
#ifndef OS2R_BLOCK
#define OS2R_BLOCK 64
#endif
#ifndef OS2R_BLOCK_WIDE
#define OS2R_BLOCK_WIDE 224
#endif
#ifndef OS2R_SORT_NARROW
#define OS2R_SORT_NARROW 0    // 1: the 64-thread blocks run the lane sort too (A/B: 2-4 % slower at every narrow-block batch size)
#endif
#ifndef OS2R_WIDE_MINB
#define OS2R_WIDE_MINB 2     // resident wide blocks per SM the build targets (A/B: 448-thread blocks, one per SM)
#endif

// Structure signatures (os2r_device.cuh) of the shipped URDFs, as os2r_model_signature reports them.
#define OS2R_SHIPPED_J23 0x40000ab8u
#define OS2R_SHIPPED_C23 0xc100u
#define OS2R_SHIPPED_J33 0x4002a6f8u
#define OS2R_SHIPPED_C33 0x14308u
#define OS2R_SHIPPED_J43 0x80a9b138u
#define OS2R_SHIPPED_C43 0x1c510u
#define OS2R_SHIPPED_J53 0xaa290138u
#define OS2R_SHIPPED_C53 0x24718u
#define OS2R_SHIPPED_J54 0xaa290138u
#define OS2R_SHIPPED_C54 0x91c616u

// builds of the step kernel
#define OS2R_BUILD_F32 0    // fp32, one env per thread: the product path
#define OS2R_BUILD_F64 2    // fp64, one env per thread (verification)

namespace os2r {

int step_block_threads(int build, int64_t n_envs, int sm_count);
template <typename T>
cudaError_t launch_step(int build, int n_dof, int n_contacts, int block, uint32_t sj, uint32_t sc, const ModelDev<T> &M,
                        const TaskDev &K, const StateDev<T> &S, const StepIO &io, StatsDev *stats, cudaStream_t stream);
template <typename T>
cudaError_t launch_reset(int n_dof, int n_contacts, const TaskDev &K, const StateDev<T> &S, const uint8_t *mask,
                         float *obs, cudaStream_t stream);
template <typename T>
cudaError_t launch_init(const TaskDev &K, const StateDev<T> &S, double nominal_gz, cudaStream_t stream);
cudaError_t prepare_step(int build, int n_dof, int n_contacts, int block, uint32_t sj, uint32_t sc);
// signature the specialised kernels of this (joints, proxies) shape were built for
bool shipped_signature(int n_dof, int n_contacts, uint32_t *sj, uint32_t *sc);
bool supported_shape(int n_dof, int n_contacts);
cudaError_t step_kernel_attributes(int build, int n_dof, int n_contacts, int block, bool damped, uint32_t sj, uint32_t sc,
                                   cudaFuncAttributes *attr, int *blocks_per_sm, int *envs_per_block);
cudaError_t read_check_counters(unsigned long long out[8], bool clear);
cudaError_t launch_fma_peak(float *out, int blocks, int iters, cudaStream_t stream);

}  // namespace os2r
