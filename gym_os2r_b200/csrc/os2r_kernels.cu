// os2r_kernels.cu — kernels of the monopod step path (sm_100a) and their launchers.
#include "os2r_kernels.h"

#include <math.h>

namespace os2r {

// ------------------------------------------------------------------------------------------------
// reset of one env, written straight to the SoA state (rare path, fp64 draws shared with the oracle)
// ------------------------------------------------------------------------------------------------
template <typename T, int N, int NC>
__device__ __noinline__ int reset_env_global(const TaskDev &K, StateDev<T> &S, int64_t e, const double *a_old,
                                             float *obs_row) {
    const int64_t NE = S.n_envs;
    const uint64_t gid = (uint64_t)(S.first_env_id + e);
    const uint32_t ep = S.episode[e] + 1u;
    S.episode[e] = ep;
    double q[N], v[N];
    const int idx = reset_pose<N>(K, S.seed, gid, ep, q);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        v[i] = 0.0;
        const T hi = (T)q[i];
        S.q_hi[i * NE + e] = hi;
        S.q_lo[i * NE + e] = (sizeof(T) == 4) ? (T)(q[i] - (double)hi) : T(0);
        S.qd[i * NE + e] = T(0);
        S.qd_lo[i * NE + e] = T(0);
    }
#pragma unroll
    for (int r = 0; r < N + 3 * NC; ++r) S.lam[r * NE + e] = T(0);
    S.reset_id[e] = idx;
    S.steps[e] = 0;
    S.ret[e] = 0.0;
    S.cls[e] = 0xFF;   // sorting hint unknown after a reset: treat every proxy as near the ground
    draw_params<T>(K, S, e, gid, ep);
    if (obs_row) {
        // the observation sees the state the device will actually integrate (hi+lo rounding of q)
        double qs[N];
#pragma unroll
        for (int i = 0; i < N; ++i) qs[i] = (double)S.q_hi[i * NE + e] + (double)S.q_lo[i * NE + e];
        double o[OS2R_MAX_OBS];
        observe<N>(K, qs, v, a_old, o);
        for (int k = 0; k < K.cfg.obs_dim; ++k) obs_row[k] = (float)o[k];
    }
    return idx;
}

template <typename T, int N, int NC>
__global__ void __launch_bounds__(OS2R_BLOCK) reset_kernel(const __grid_constant__ TaskDev K, StateDev<T> S,
                                                           const uint8_t *mask, float *obs) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S.n_envs) return;
    if (mask && !mask[e]) return;
    const double a_old[2] = {(double)S.a_prev[e], (double)S.a_prev[S.n_envs + e]};
    reset_env_global<T, N, NC>(K, S, e, a_old, obs ? obs + e * K.cfg.obs_dim : nullptr);
}

// nominal parameters + the once-per-env gravity draw (GazeboEnvRandomizer.__init__ -> randomize_physics)
template <typename T>
__global__ void __launch_bounds__(OS2R_BLOCK) init_kernel(const __grid_constant__ TaskDev K, StateDev<T> S, double nominal_gz) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t NE = S.n_envs;
    if (e >= NE) return;
    const uint64_t gid = (uint64_t)(S.first_env_id + e);
    for (int i = 0; i < K.n_dof; ++i) {
        S.q_hi[i * NE + e] = T(0); S.q_lo[i * NE + e] = T(0); S.qd[i * NE + e] = T(0); S.qd_lo[i * NE + e] = T(0);
        S.mass_scale[i * NE + e] = T(1);
        S.damping[i * NE + e] = (T)K.nominal_damping[i];
        S.friction[i * NE + e] = (T)K.nominal_friction[i];
    }
    for (int c = 0; c < K.n_contacts; ++c) S.mu[c * NE + e] = (T)K.nominal_mu[c];
    for (int r = 0; r < K.n_dof + 3 * K.n_contacts; ++r) S.lam[r * NE + e] = T(0);
    S.a_prev[e] = T(0); S.a_prev[NE + e] = T(0);
    double gz = nominal_gz;
    if (K.cfg.randomize_gravity) {
        double z[2];
        rng_normal2(S.seed, gid, OS2R_EPISODE_GRAVITY, 0, z);
        gz = K.cfg.grav_mean + K.cfg.grav_std * z[0];
    }
    S.gravity_z[e] = (T)gz;
    S.steps[e] = 0; S.episode[e] = 0; S.reset_id[e] = 0; S.ret[e] = 0.0; S.cls[e] = 0xFF;
}

// ------------------------------------------------------------------------------------------------
// lane sorting: which env of the block's window each thread steps
// ------------------------------------------------------------------------------------------------
// A warp pays for the union of its lanes' constraint rows (each contact proxy costs ~780 instructions per
// physics iteration as soon as ONE lane touches the ground), so the block regroups its envs by the set of
// proxies that were near the ground after the previous step: envs lying on the ground share a warp, hopping
// envs share the next ones, airborne envs fill the rest and never execute contact rows. Stable counting sort
// over 2^NC classes (+1 for slots past the end of the batch), ballot/match based, once per env step.
// The hint only decides the thread <-> env pairing; every per-env result is independent of it.
template <int BLOCK, int NKEYS>
__device__ __forceinline__ int sorted_source(int key, int *scratch) {
    constexpr int W = BLOCK / 32;
    constexpr int CNT = NKEYS * W;                 // counters, ordered (heavier class first, then warp)
    constexpr int PER = (CNT + 31) / 32;           // counters scanned by each lane of warp 0
    constexpr int CPAD = PER * 32;
    int *cnt = scratch;              // [CPAD] counts, then exclusive offsets
    int *perm = scratch + CPAD;      // [BLOCK]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < CPAD; k += BLOCK) cnt[k] = 0;
    __syncthreads();
    const unsigned same = __match_any_sync(0xffffffffu, key);
    const int p = (NKEYS - 1 - key) * W + warp;
    if (lane == __ffs(same) - 1) cnt[p] = __popc(same);
    __syncthreads();
    if (warp == 0) {
        int v[PER], s = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) { v[k] = cnt[PER * lane + k]; s += v[k]; }
        const int own = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        int run = s - own;
#pragma unroll
        for (int k = 0; k < PER; ++k) { cnt[PER * lane + k] = run; run += v[k]; }
    }
    __syncthreads();
    perm[cnt[p] + __popc(same & ((1u << lane) - 1u))] = tid;
    __syncthreads();
    const int src = perm[tid];
    __syncthreads();                 // the scratch area is reused for the cold slots
    return src;
}

// ------------------------------------------------------------------------------------------------
// the fused step kernel: substeps x physics + observation + reward + done + auto-reset
// ------------------------------------------------------------------------------------------------
// LONE = the build for batches of at most four 2-warp blocks per SM (every batch below the wide-block threshold): no
// occupancy target, so ptxas takes ~200-225 registers instead of 128 and schedules for instruction-level parallelism
// (with one or two warps per scheduler a warp is bound by its own dependency chains): 4-12 % lower step latency.
template <typename T, int N, int NC, int BLOCK, bool DAMPED, bool LONE>
__global__ void __launch_bounds__(BLOCK, (sizeof(T) == 4 && !LONE ? OS2R_RESIDENT_THREADS / BLOCK : 1))
step_kernel(const __grid_constant__ ModelDev<T> M, const __grid_constant__ TaskDev K, StateDev<T> S,
            const __grid_constant__ StepIO IO, StatsDev *stats) {
    const float *__restrict__ actions = IO.actions;
    float *__restrict__ obs = IO.obs;
    float *__restrict__ reward = IO.reward;
    uint8_t *__restrict__ done = IO.done;
    float *__restrict__ term_obs = IO.term_obs;
    int32_t *__restrict__ info = IO.info;
    static_assert(NC <= 3, "class byte: one bit per contact proxy, 2^NC + 1 sort keys");
    const int64_t NE = S.n_envs;
    using SL = ColdSlots<N, NC>;
    constexpr int ROWS = SL::ROWS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // Thread t steps env (block window start + src). Slots past the end of the batch (last block only) sort last
    // and shadow the last env instead of exiting, so that the block-wide barriers inside the physics loop stay
    // legal; they leave before anything is written.
    const int64_t window = (int64_t)blockIdx.x * BLOCK;
    const bool nominal_valid = window + threadIdx.x < NE;
    const int key = nominal_valid ? 1 + (int)(__ldcg(S.cls + window + threadIdx.x) & ((1u << NC) - 1u)) : 0;
    // While the class byte travels and the block sorts, pull the window's state lines towards the L2: which env a
    // thread will step is not known yet, but it is one of this window's, so the DRAM latency of the prologue loads
    // overlaps the sort instead of following it.
    if (nominal_valid) {
        const int64_t en = window + threadIdx.x;
        auto pf = [](const void *ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); };
#pragma unroll
        for (int i = 0; i < N; ++i) {
            pf(S.q_hi + i * NE + en); pf(S.qd + i * NE + en); pf(S.q_lo + i * NE + en); pf(S.qd_lo + i * NE + en);
            pf(S.mass_scale + i * NE + en); pf(S.damping + i * NE + en); pf(S.friction + i * NE + en);
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) pf(S.lam + r * NE + en);
#pragma unroll
        for (int c = 0; c < NC; ++c) pf(S.mu + c * NE + en);
        pf(S.gravity_z + en); pf(S.a_prev + en); pf(S.a_prev + NE + en); pf(S.steps + en); pf(S.reset_id + en);
        pf(S.ret + en); pf(reinterpret_cast<const float2 *>(IO.actions) + en);
    }
    const int src = sorted_source<BLOCK, (1 << NC) + 1>(key, reinterpret_cast<int *>(smem_raw));
    const int64_t e_raw = window + src;
    const bool valid = e_raw < NE;
    const int64_t e = valid ? e_raw : NE - 1;
    const Cold<T, BLOCK> C{reinterpret_cast<T *>(smem_raw) + threadIdx.x};

    EnvRegs<T, N> E;
    // ---- prologue: ALL global loads are issued back to back (explicit ld.global, so the compiler may hoist
    //      them above the shared-memory stores that follow; generic loads were serialised LD -> STS -> LD ...,
    //      one DRAM round trip each) and only then scattered into registers / shared memory.
    T ld_qlo[N], ld_vlo[N], ld_mass[N], ld_damp[N], ld_fric[N], ld_lam[ROWS], ld_mu[NC];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        E.q_hi[i] = __ldcg(S.q_hi + i * NE + e);
        E.v[i] = __ldcg(S.qd + i * NE + e);
        ld_qlo[i] = __ldcg(S.q_lo + i * NE + e);
        ld_vlo[i] = __ldcg(S.qd_lo + i * NE + e);
        ld_mass[i] = __ldcg(S.mass_scale + i * NE + e);
        ld_damp[i] = __ldcg(S.damping + i * NE + e);
        ld_fric[i] = __ldcg(S.friction + i * NE + e);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) ld_lam[r] = __ldcg(S.lam + r * NE + e);
#pragma unroll
    for (int c = 0; c < NC; ++c) ld_mu[c] = __ldcg(S.mu + c * NE + e);
    E.gz = __ldcg(S.gravity_z + e);
    const T a_old0 = __ldcg(S.a_prev + e), a_old1 = __ldcg(S.a_prev + NE + e);
    const int steps_in = __ldcg(S.steps + e);
    int reset_idx = __ldcg(S.reset_id + e);
    const double ret_in = __ldcg(S.ret + e);
    float2 act = __ldcg(reinterpret_cast<const float2 *>(actions) + e);
    // A non-finite action (the reference rejects it through `assert action_space.contains`, tasks/monopod.py:218)
    // applies no torque and takes the non-finite path of the epilogue: defined outputs, forced reset, counted.
    // fminf / fmaxf would silently turn a NaN into a full negative torque.
    const bool act_finite = isfinite(act.x) && isfinite(act.y);
    if (!act_finite) { act.x = 0.0f; act.y = 0.0f; }
    // ScenarIO clips force targets to +-max force (tasks/monopod.py:313-316)
    act.x = fminf(1.0f, fmaxf(-1.0f, act.x));
    act.y = fminf(1.0f, fmaxf(-1.0f, act.y));
#pragma unroll
    for (int i = 0; i < N; ++i) {
        C(SL::QLO + i) = ld_qlo[i];
        C(SL::VLO + i) = ld_vlo[i];
        C(SL::MASS + i) = ld_mass[i];
        C(SL::DAMP + i) = ld_damp[i];
        C(SL::FRIC + i) = ld_fric[i] * M.dt;
        T tau = T(0);
        if (i == M.hip_dof) tau = M.max_torque[0] * (T)act.x;
        if (i == M.knee_dof) tau = M.max_torque[1] * (T)act.y;
        C(SL::TAU + i) = tau;
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) C(SL::LAM + r) = ld_lam[r];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        C(SL::MU + c) = ld_mu[c];
        C(SL::CX + 3 * c + 2) = T(0);   // defined even when substeps == 0 (reads as "near")
    }
    C(SL::AOLD) = a_old0;
    C(SL::AOLD + 1) = a_old1;
    C(SL::MISC) = __int_as_float_t<T>(steps_in);
    C(SL::MISC + 1) = __int_as_float_t<T>(__double2loint(ret_in));
    C(SL::MISC + 2) = __int_as_float_t<T>(__double2hiint(ret_in));

#pragma unroll 1
    for (int s = 0; s < M.substeps; ++s) {
        // Keep the block's warps in phase: all warps of an SM run the same ~41 KB loop body, and warps that drift
        // apart thrash the instruction caches (measured: -3..4 % step time with the two barriers per iteration;
        // a rolled forward pass that fits the 32 KB cache still lost 12 % without them and 20 % overall to its
        // extra instructions and spills — DESIGN.md section 9).
        __syncthreads();
        physics_iteration<T, N, NC, DAMPED, Cold<T, BLOCK>>(M, E, C);
    }
    if (!valid) return;

    // ---- epilogue (fp64, once per env step) --------------------------------------------------------
    double q[N], v[N];
    bool finite = act_finite;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        q[i] = (double)E.q_hi[i] + (double)C(SL::QLO + i);
        v[i] = (double)E.v[i] + (double)C(SL::VLO + i);
        finite = finite && isfinite(q[i]) && isfinite(v[i]);
    }
    if (!finite) {   // nothing non-finite leaves the kernel: the step reports a zero observation and reward
#pragma unroll
        for (int i = 0; i < N; ++i) { q[i] = 0.0; v[i] = 0.0; }
    }
    const double a0[2] = {(double)act.x, (double)act.y};
    const double a_old[2] = {(double)C(SL::AOLD), (double)C(SL::AOLD + 1)};
    double o[OS2R_MAX_OBS];
    const bool task_done = observe<N>(K, q, v, a_old, o);
    const double r = finite ? reward_fn(K.cfg, o, a0, a_old) : 0.0;
    const int D = K.cfg.obs_dim;
    int cause = (task_done && finite) ? 1 : 0;
    if (!finite) cause = 4;
    const int steps = __float_as_int_t<T>(C(SL::MISC)) + 1;
    const double ret = __hiloint2double(__float_as_int_t<T>(C(SL::MISC + 2)), __float_as_int_t<T>(C(SL::MISC + 1))) + r;
    if (K.cfg.max_episode_steps > 0 && steps >= K.cfg.max_episode_steps) cause |= 2;

    reward[e] = (float)r;
    done[e] = cause != 0;
    if (term_obs) for (int k = 0; k < D; ++k) term_obs[e * D + k] = (float)o[k];
    S.a_prev[e] = (T)act.x;
    S.a_prev[NE + e] = (T)act.y;

    if (cause && IO.term_count) {
        const int k = atomicAdd(IO.term_count, 1);
        if (k < IO.term_cap) {
            int32_t *rec = IO.term_records + (int64_t)k * (D + 2);
            rec[0] = (int32_t)e;
            rec[1] = cause;
            for (int c = 0; c < D; ++c) rec[2 + c] = __float_as_int((float)o[c]);
        }
    }
    if (cause) {
        atomicAdd(&stats->episodes, 1ull);
        if (cause & 1) atomicAdd(&stats->done_task, 1ull);
        if (cause & 2) atomicAdd(&stats->done_timelimit, 1ull);
        if (cause & 4) atomicAdd(&stats->nonfinite_resets, 1ull);
        atomicAdd(&stats->sum_return, ret);
        atomicAdd(&stats->sum_length, (double)steps);
    }
    if (cause && (K.cfg.auto_reset || !finite)) {
        reset_idx = reset_env_global<T, N, NC>(K, S, e, a_old, obs + e * D);   // also marks the env "unknown" (all near)
    } else {
        // next step's sorting hint: which proxies ended within sort_margin of the ground
        unsigned cls = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (!(C(SL::CX + 3 * c + 2) - M.contact_radius[c] >= M.sort_margin)) cls |= 1u << c;
        S.cls[e] = (uint8_t)cls;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            S.q_hi[i * NE + e] = E.q_hi[i];
            S.q_lo[i * NE + e] = C(SL::QLO + i);
            S.qd[i * NE + e] = E.v[i];
            S.qd_lo[i * NE + e] = C(SL::VLO + i);
        }
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr) S.lam[rr * NE + e] = C(SL::LAM + rr);
        S.steps[e] = steps;
        S.ret[e] = ret;
        for (int k = 0; k < D; ++k) obs[e * D + k] = (float)o[k];
    }
    if (info) {
        info[2 * e] = reset_idx;
        info[2 * e + 1] = cause;
    }
    if (IO.reset_id8) IO.reset_id8[e] = (uint8_t)reset_idx;
}

// ------------------------------------------------------------------------------------------------
// fp32 FMA-pipe peak microbenchmark (roofline denominator; MEASURED_PEAKS.json has no fp32 entry)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fma_peak_kernel(float *out, int iters, float seed) {
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float m = 0.999f, c = 1e-3f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------------------------------------
// launchers (dispatch on n_dof; all shipped models carry 3 contact proxies)
// ------------------------------------------------------------------------------------------------
static inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

#define OS2R_DISPATCH_N(n_dof, CALL)                 \
    switch (n_dof) {                                 \
    case 2: { constexpr int N_ = 2; CALL; } break;   \
    case 3: { constexpr int N_ = 3; CALL; } break;   \
    case 4: { constexpr int N_ = 4; CALL; } break;   \
    case 5: { constexpr int N_ = 5; CALL; } break;   \
    default: return cudaErrorInvalidValue;           \
    }

template <typename T, int N, int BLOCK>
static constexpr size_t step_smem_bytes() {
    size_t cold = (size_t)ColdSlots<N, OS2R_NC>::COUNT * BLOCK * sizeof(T), sort = (size_t)(512 + BLOCK) * sizeof(int);
    return cold > sort ? cold : sort;
}

// Wide blocks (7 warps) give the lane sort enough envs to fill whole warps with one class; they need at least one
// block per SM to pay off. Small batches and the fp64 verification build keep 2-warp blocks (more SMs busy).
template <typename T>
int step_block_threads(int64_t n_envs, int sm_count) {
    if (sizeof(T) == 4 && n_envs >= (int64_t)sm_count * OS2R_BLOCK_WIDE) return OS2R_BLOCK_WIDE;
    return OS2R_BLOCK;
}

template <typename T, int N, int BLOCK, bool DAMPED>
static cudaError_t launch_step_nd(const ModelDev<T> &M, const TaskDev &K, const StateDev<T> &S, const StepIO &io, StatsDev *stats,
                                  cudaStream_t stream, bool lone) {
    constexpr size_t smem = step_smem_bytes<T, N, BLOCK>();
    if constexpr (sizeof(T) == 4 && BLOCK == OS2R_BLOCK) {
        if (lone) {
            step_kernel<T, N, OS2R_NC, BLOCK, DAMPED, true><<<grid_for(S.n_envs, BLOCK), BLOCK, smem, stream>>>(M, K, S, io, stats);
            return cudaGetLastError();
        }
    }
    // blocks that need more than 48 KB of dynamic shared memory were opted in by prepare_step (once per handle,
    // on the handle's device: the attribute is per device, a process can hold handles on several GPUs)
    step_kernel<T, N, OS2R_NC, BLOCK, DAMPED, false><<<grid_for(S.n_envs, BLOCK), BLOCK, smem, stream>>>(M, K, S, io, stats);
    return cudaGetLastError();
}
template <typename T, int N, int BLOCK>
static cudaError_t launch_step_n(const ModelDev<T> &M, const TaskDev &K, const StateDev<T> &S, const StepIO &io, StatsDev *stats,
                                 cudaStream_t stream, bool lone) {
    // the fp64 verification build keeps one (damped) instantiation; with zero damping its second factor equals the first
    if (sizeof(T) == 8 || M.any_damping) return launch_step_nd<T, N, BLOCK, true>(M, K, S, io, stats, stream, lone);
    if constexpr (sizeof(T) == 4) return launch_step_nd<T, N, BLOCK, false>(M, K, S, io, stats, stream, lone);
    return cudaErrorInvalidValue;
}

template <typename T>
cudaError_t launch_step(int n_dof, int n_contacts, int block, bool lone, const ModelDev<T> &M, const TaskDev &K,
                        const StateDev<T> &S, const StepIO &io, StatsDev *stats, cudaStream_t stream) {
    if (n_contacts != OS2R_NC) return cudaErrorInvalidValue;
    if (sizeof(T) == 4 && block == OS2R_BLOCK_WIDE) {
        if constexpr (sizeof(T) == 4) {
            OS2R_DISPATCH_N(n_dof, return (launch_step_n<T, N_, OS2R_BLOCK_WIDE>(M, K, S, io, stats, stream, false)));
        }
    }
    if (block != OS2R_BLOCK) return cudaErrorInvalidValue;
    OS2R_DISPATCH_N(n_dof, return (launch_step_n<T, N_, OS2R_BLOCK>(M, K, S, io, stats, stream, lone)));
    return cudaErrorInvalidValue;
}

template <typename T>
cudaError_t launch_reset(int n_dof, int n_contacts, const TaskDev &K, const StateDev<T> &S, const uint8_t *mask,
                         float *obs, cudaStream_t stream) {
    if (n_contacts != OS2R_NC) return cudaErrorInvalidValue;
    OS2R_DISPATCH_N(n_dof, (reset_kernel<T, N_, OS2R_NC><<<grid_for(S.n_envs, OS2R_BLOCK), OS2R_BLOCK, 0, stream>>>(K, S, mask, obs)));
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_init(const TaskDev &K, const StateDev<T> &S, double nominal_gz, cudaStream_t stream) {
    init_kernel<T><<<grid_for(S.n_envs, OS2R_BLOCK), OS2R_BLOCK, 0, stream>>>(K, S, nominal_gz);
    return cudaGetLastError();
}

template <typename T, int N, int BLOCK, bool DAMPED>
static cudaError_t step_attr_nd(cudaFuncAttributes *attr, int *blocks_per_sm) {
    constexpr size_t smem = step_smem_bytes<T, N, BLOCK>();
    cudaError_t e = cudaFuncGetAttributes(attr, step_kernel<T, N, OS2R_NC, BLOCK, DAMPED, false>);
    if (e != cudaSuccess) return e;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(step_kernel<T, N, OS2R_NC, BLOCK, DAMPED, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, step_kernel<T, N, OS2R_NC, BLOCK, DAMPED, false>, BLOCK, smem);
}
template <typename T, int N, int BLOCK>
static cudaError_t step_attr_n(bool damped, cudaFuncAttributes *attr, int *blocks_per_sm) {
    if (sizeof(T) == 8 || damped) return step_attr_nd<T, N, BLOCK, true>(attr, blocks_per_sm);
    if constexpr (sizeof(T) == 4) return step_attr_nd<T, N, BLOCK, false>(attr, blocks_per_sm);
    return cudaErrorInvalidValue;
}

template <typename T>
cudaError_t step_kernel_attributes(int n_dof, int block, bool damped, cudaFuncAttributes *attr, int *blocks_per_sm) {
    if (sizeof(T) == 4 && block == OS2R_BLOCK_WIDE) {
        if constexpr (sizeof(T) == 4) {
            OS2R_DISPATCH_N(n_dof, return (step_attr_n<T, N_, OS2R_BLOCK_WIDE>(damped, attr, blocks_per_sm)));
        }
    }
    OS2R_DISPATCH_N(n_dof, return (step_attr_n<T, N_, OS2R_BLOCK>(damped, attr, blocks_per_sm)));
    return cudaErrorInvalidValue;
}

template <typename T, int N, int BLOCK, bool DAMPED>
static cudaError_t prepare_nd() {
    constexpr size_t smem = step_smem_bytes<T, N, BLOCK>();
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(step_kernel<T, N, OS2R_NC, BLOCK, DAMPED, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <typename T, int N, int BLOCK>
static cudaError_t prepare_n() {
    cudaError_t e = prepare_nd<T, N, BLOCK, true>();
    if (e != cudaSuccess) return e;
    if constexpr (sizeof(T) == 4) return prepare_nd<T, N, BLOCK, false>();
    return cudaSuccess;
}
// Opt the step kernels this handle can launch (damped and undamped build) into their dynamic shared memory size on the
// CURRENT device. Called by os2r_create under its device guard.
template <typename T>
cudaError_t prepare_step(int n_dof, int block) {
    if (sizeof(T) == 4 && block == OS2R_BLOCK_WIDE) {
        if constexpr (sizeof(T) == 4) {
            OS2R_DISPATCH_N(n_dof, return (prepare_n<T, N_, OS2R_BLOCK_WIDE>()));
        }
    }
    OS2R_DISPATCH_N(n_dof, return (prepare_n<T, N_, OS2R_BLOCK>()));
    return cudaErrorInvalidValue;
}

cudaError_t launch_fma_peak(float *out, int blocks, int iters, cudaStream_t stream) {
    fma_peak_kernel<<<blocks, 256, 0, stream>>>(out, iters, 1.0f);
    return cudaGetLastError();
}

#define OS2R_INSTANTIATE(T)                                                                                   \
    template int step_block_threads<T>(int64_t, int);                                                         \
    template cudaError_t launch_step<T>(int, int, int, bool, const ModelDev<T> &, const TaskDev &, const StateDev<T> &, \
                                        const StepIO &, StatsDev *, cudaStream_t);                            \
    template cudaError_t launch_reset<T>(int, int, const TaskDev &, const StateDev<T> &, const uint8_t *,     \
                                         float *, cudaStream_t);                                              \
    template cudaError_t launch_init<T>(const TaskDev &, const StateDev<T> &, double, cudaStream_t);          \
    template cudaError_t step_kernel_attributes<T>(int, int, bool, cudaFuncAttributes *, int *);            \
    template cudaError_t prepare_step<T>(int, int);
OS2R_INSTANTIATE(float)
OS2R_INSTANTIATE(double)

}  // namespace os2r
