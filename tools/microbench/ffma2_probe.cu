// ffma2_probe.cu — issue rate and dependent-issue latency of scalar FFMA (register / constant-bank / immediate
// operand forms) against the packed `fma.rn.f32x2` (SASS FFMA2, sm_100+), for the warps-per-scheduler shapes the step
// kernel runs at (14 warps/SM one env per thread, 7 warps/SM two envs per thread).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo_of(unsigned long long v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

struct Consts { float m[8], c[8]; };

// MODE 0: FFMA, all three operands in registers (m, c loaded from global: the compiler cannot fold them)
// MODE 1: FFMA with constant-bank operands (kernel parameters)
// MODE 2: FFMA2 packed, registers
// MODE 3: FFMA with immediates
template <int MODE, int ILP>
__global__ void __launch_bounds__(256) rate_kernel(float *out, const float *in, int iters, const __grid_constant__ Consts K) {
    float a[ILP];
    unsigned long long p[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { a[k] = threadIdx.x + k; p[k] = pack(a[k], a[k] + 0.5f); }
    const float mr = in[0], cr = in[1];
    const float mr2 = in[2], cr2 = in[3];
    const unsigned long long mp = pack(mr, mr2), cp = pack(cr, cr2);
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                if (MODE == 0) a[k] = fmaf(a[k], (u & 1) ? mr : mr2, (u & 2) ? cr : cr2);
                if (MODE == 1) a[k] = fmaf(a[k], K.m[u & 7], K.c[k & 7]);
                if (MODE == 2) p[k] = fma2(p[k], mp, cp);
                if (MODE == 3) a[k] = fmaf(a[k], 0.999f, 1e-3f);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += (MODE == 2) ? lo_of(p[k]) : a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MODE 4 / 5: scalar FFMA / packed FFMA2 whose three operands are DISTINCT registers that change from instruction to
// instruction (no operand-reuse cache hits, register-bank conflicts as in real code): 8 independent accumulators.
template <int MODE>
__global__ void __launch_bounds__(256) mix_kernel(float *out, const float *in, int iters) {
    float a[8], b[8], c[8];
    unsigned long long pa[8], pb[8], pc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = in[k] + threadIdx.x * 1e-6f; b[k] = in[8 + k]; c[k] = in[16 + k];
        pa[k] = pack(a[k], a[k] * 0.5f); pb[k] = pack(b[k], b[k] * 0.5f); pc[k] = pack(c[k], c[k] * 0.5f);
    }
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 4) c[k] = fmaf(a[(k + u) & 7], b[(k + 3 * u + 1) & 7], c[k]);
                else pc[k] = fma2(pa[(k + u) & 7], pb[(k + 3 * u + 1) & 7], pc[k]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += (MODE == 5) ? lo_of(pc[k]) : c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run_mix(const char *name, float *out, const float *in, int sms, double ghz, int threads) {
    const int iters = 20000;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(t0);
        mix_kernel<MODE><<<sms, threads>>>(out, in, iters);
        cudaEventRecord(t1);
        cudaEventSynchronize(t1);
        float ms; cudaEventElapsedTime(&ms, t0, t1);
        if (rep && ms < best) best = ms;
    }
    const double cycles = best * 1e-3 * ghz * 1e9, instr_per_warp = 16.0 * 8 * iters, warps = threads / 32.0;
    const double fma_per_instr = MODE == 5 ? 64.0 : 32.0;
    printf("%-28s warps/SM %4.1f: cycles per warp-instr %6.3f  SM IPC %5.3f  FMA/clk/SM %6.1f\n", name, warps,
           cycles / instr_per_warp, instr_per_warp * warps / cycles, instr_per_warp * warps * fma_per_instr / cycles);
}

template <int MODE, int ILP>
void run(const char *name, float *out, const float *in, int sms, double ghz, int threads) {
    Consts K;
    for (int i = 0; i < 8; ++i) { K.m[i] = 0.999f - 1e-4f * i; K.c[i] = 1e-3f + 1e-5f * i; }
    const int iters = 20000;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(t0);
        rate_kernel<MODE, ILP><<<sms, threads>>>(out, in, iters, K);
        cudaEventRecord(t1);
        cudaEventSynchronize(t1);
        float ms; cudaEventElapsedTime(&ms, t0, t1);
        if (rep && ms < best) best = ms;
    }
    const double cycles = best * 1e-3 * ghz * 1e9;
    const double instr_per_warp = 16.0 * ILP * iters;
    const double warps = threads / 32.0;
    const double fma_per_instr = MODE == 2 ? 64.0 : 32.0;
    printf("%-18s ILP %d  warps/SM %4.1f: cycles per warp-instr %6.3f  SM IPC %5.3f  FMA/clk/SM %6.1f  (%.1f TFLOP/s)\n", name, ILP,
           warps, cycles / instr_per_warp, instr_per_warp * warps / cycles, instr_per_warp * warps * fma_per_instr / cycles,
           instr_per_warp * warps * fma_per_instr * 2 * sms / (best * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const double ghz = p.clockRate / 1e6;
    printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    float *out, *in; cudaMalloc(&out, 1 << 24); cudaMalloc(&in, 256);
    const float h[4] = {0.999f, 1e-3f, 0.998f, 2e-3f};
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const int sms = p.multiProcessorCount;
    float hin[24];
    for (int i = 0; i < 24; ++i) hin[i] = 0.5f + 0.01f * i;
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    for (int threads : {32, 64, 128, 224, 256}) {
        run_mix<4>("FFMA distinct operands", out, in, sms, ghz, threads);
        run_mix<5>("FFMA2 distinct operands", out, in, sms, ghz, threads);
    }
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const int shapes[] = {32, 224};
    for (int threads : shapes) {
        run<0, 1>("FFMA reg", out, in, sms, ghz, threads);
        run<0, 2>("FFMA reg", out, in, sms, ghz, threads);
        run<0, 4>("FFMA reg", out, in, sms, ghz, threads);
        run<0, 8>("FFMA reg", out, in, sms, ghz, threads);
        run<1, 1>("FFMA const-bank", out, in, sms, ghz, threads);
        run<1, 4>("FFMA const-bank", out, in, sms, ghz, threads);
        run<1, 8>("FFMA const-bank", out, in, sms, ghz, threads);
        run<3, 8>("FFMA imm", out, in, sms, ghz, threads);
        run<2, 1>("FFMA2 packed", out, in, sms, ghz, threads);
        run<2, 2>("FFMA2 packed", out, in, sms, ghz, threads);
        run<2, 4>("FFMA2 packed", out, in, sms, ghz, threads);
        run<2, 8>("FFMA2 packed", out, in, sms, ghz, threads);
    }
    return 0;
}
