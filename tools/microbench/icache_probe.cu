// icache_probe.cu — how fast does ONE warp (or k warps per SM) run straight-line FFMA code whose loop body is
// larger than the instruction caches?  Body = BODY independent-ish FFMAs (8 accumulators, so no dependency
// stalls), executed `iters` times.  Prints cycles per warp-instruction for several body sizes / warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache_probe icache_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int BODY>
__global__ void __launch_bounds__(512) body_kernel(float *out, int iters, float m, float c) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < BODY / 8; ++u) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int BODY>
void run(float *out, int sms, double ghz) {
    const int threads_list[] = {32, 128, 448};
    for (int threads : threads_list) {
        const int iters = 4 * 1024 * 1024 / BODY / (threads >= 448 ? 4 : 1);
        cudaEvent_t t0, t1;
        cudaEventCreate(&t0); cudaEventCreate(&t1);
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(t0);
            body_kernel<BODY><<<sms, threads>>>(out, iters, 0.999f, 1e-3f);
            cudaEventRecord(t1);
            cudaEventSynchronize(t1);
            float ms; cudaEventElapsedTime(&ms, t0, t1);
            if (rep && ms < best) best = ms;
        }
        const double cycles = best * 1e-3 * ghz * 1e9;
        const double instr_per_warp = (double)BODY * iters;
        const int warps_per_sched = (threads / 32 + 3) / 4;
        printf("body %5d instr (%3d KB)  warps/SM %2d: %8.3f ms  cycles per warp-instr %.3f  per-scheduler IPC %.3f\n", BODY,
               BODY * 16 / 1024, threads / 32, best, cycles / instr_per_warp, instr_per_warp * (threads / 32) / 4.0 / cycles * (threads >= 128 ? 1 : 4));
        (void)warps_per_sched;
    }
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const double ghz = p.clockRate / 1e6;
    printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    float *out; cudaMalloc(&out, 1 << 24);
    run<512>(out, p.multiProcessorCount, ghz);
    run<1024>(out, p.multiProcessorCount, ghz);
    run<1536>(out, p.multiProcessorCount, ghz);
    run<2048>(out, p.multiProcessorCount, ghz);
    run<2560>(out, p.multiProcessorCount, ghz);
    run<3072>(out, p.multiProcessorCount, ghz);
    run<4096>(out, p.multiProcessorCount, ghz);
    run<6144>(out, p.multiProcessorCount, ghz);
    return 0;
}
