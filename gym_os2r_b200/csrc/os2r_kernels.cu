// os2r_kernels.cu — kernels of the monopod step path (sm_100a) and their launchers.
#include "os2r_kernels.h"

#include <math.h>

namespace os2r {

// Checked build (-DOS2R_CHECKED, `make libos2r_checked.so`, tools/sanitize.sh): compute-sanitizer is closed on this
// GPU pool, so the step kernel carries its own bounds / permutation checks; violations are COUNTED (a trap would take
// the context down) and read back through os2r_debug_counters.
//   [0] sorted source slot outside the block's window      [1] lane sort did not produce a permutation
//   [2] env index outside the batch                        [3] terminal-record list overflowed its capacity
//   [4] cold-slot guard word overwritten                   [7] 1 = this library was built with the checks
__device__ unsigned long long g_check[8];
#ifdef OS2R_CHECKED
#define OS2R_CHECK(cond, k) do { if (!(cond)) atomicAdd(&g_check[k], 1ull); } while (0)
#else
#define OS2R_CHECK(cond, k) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------
// reset of one env, written straight to the SoA state (rare path, fp64 draws shared with the oracle)
// ------------------------------------------------------------------------------------------------
// `draw(k)` yields the k-th uniform of episode `ep` of this env (os2r_device.cuh: DrawsPhilox / DrawsTable).
template <typename T, int N, int NC, typename D>
__device__ __noinline__ int reset_env_global(const TaskDev &K, StateDev<T> &S, int64_t e, const double *a_old,
                                             float *obs_row, uint32_t ep, const D &draw) {
    const int64_t NE = S.n_envs;
    S.episode[e] = ep;
    double q[N], v[N];
    const int idx = reset_pose<N>(K, draw, q);
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
    draw_params<T>(K, S, e, ep, draw);
    if (obs_row) {
        // the observation sees the state the device will actually integrate (hi+lo rounding of q)
        double qs[N];
#pragma unroll
        for (int i = 0; i < N; ++i) qs[i] = (double)S.q_hi[i * NE + e] + (double)S.q_lo[i * NE + e];
        double o[OS2R_MAX_OBS];
        observe<N, sizeof(T) == 4>(K, qs, v, a_old, o);
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
    const DrawsPhilox draw{S.seed, (uint64_t)(S.first_env_id + e), S.episode[e] + 1u};
    reset_env_global<T, N, NC>(K, S, e, a_old, obs ? obs + e * K.cfg.obs_dim : nullptr, draw.ep, draw);
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
// lane sorting: which env(s) of the block's window each thread steps
// ------------------------------------------------------------------------------------------------
// A warp pays for the union of its lanes' constraint rows (each contact proxy costs ~780 instructions per
// physics iteration as soon as ONE lane touches the ground), so the block regroups its envs by the set of
// proxies that were near the ground after the previous step: envs lying on the ground share a warp (and, in the pair
// build, a thread), hopping envs share the next ones, airborne envs fill the rest and never execute contact rows.
// Stable counting sort of the window's BLOCK * LANES envs over 2^NC classes (+1 for slots past the end of the batch),
// ballot/match based, once per env step: window slot h * BLOCK + tid carries key[h]; thread t then steps the envs at
// sorted positions LANES * t .. LANES * t + LANES - 1. The hint only decides the thread <-> env pairing; every per-env
// result is independent of it.
template <int BLOCK, int LANES, int NKEYS>
__device__ __forceinline__ void sorted_sources(const int (&key)[LANES], int (&src)[LANES], int *scratch) {
    constexpr int W = BLOCK / 32;
    constexpr int WL = W * LANES;
    constexpr int CNT = NKEYS * WL;                // counters, ordered (heavier class first, then half, then warp)
    constexpr int PER = (CNT + 31) / 32;           // counters scanned by each lane of warp 0
    constexpr int CPAD = PER * 32;
    int *cnt = scratch;              // [CPAD] counts, then exclusive offsets
    int *perm = scratch + CPAD;      // [BLOCK * LANES]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < CPAD; k += BLOCK) cnt[k] = 0;
    __syncthreads();
    unsigned same[LANES];
    int p[LANES];
#pragma unroll
    for (int h = 0; h < LANES; ++h) {
        same[h] = __match_any_sync(0xffffffffu, key[h]);
        p[h] = (NKEYS - 1 - key[h]) * WL + h * W + warp;
        if (lane == __ffs(same[h]) - 1) cnt[p[h]] = __popc(same[h]);
    }
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
#pragma unroll
    for (int h = 0; h < LANES; ++h) perm[cnt[p[h]] + __popc(same[h] & ((1u << lane) - 1u))] = h * BLOCK + tid;
    __syncthreads();
#pragma unroll
    for (int h = 0; h < LANES; ++h) src[h] = perm[LANES * tid + h];
    __syncthreads();                 // the scratch area is reused for the cold slots
}

// gather one value per env of the thread into a V
template <typename V>
__device__ __forceinline__ V gather(const typename VT<V>::S *p, const int64_t (&e)[VT<V>::LANES]) {
    if constexpr (VT<V>::LANES == 2) return V(__ldcg(p + e[0]), __ldcg(p + e[1]));
    else return __ldcg(p + e[0]);
}
template <typename V>
__device__ __forceinline__ V from_halves(const float (&x)[VT<V>::LANES]) {
    if constexpr (VT<V>::LANES == 2) return V(x[0], x[1]);
    else return (V)x[0];
}

// ------------------------------------------------------------------------------------------------
// the fused step kernel: substeps x physics + observation + reward + done + auto-reset
// ------------------------------------------------------------------------------------------------
// V = float : the product build, one env per thread. Batches of >= one 224-thread block per SM run 7-warp blocks, two
//     per SM (MINB = 2: 128 registers per thread; 65 536 envs = 293 blocks = ONE wave of 14 warps per SM); smaller
//     batches run 2-warp blocks without an occupancy target (MINB = 1: ptxas takes ~200-225 registers and schedules for
//     instruction-level parallelism — with one or two warps per scheduler a warp is bound by its own dependency chains);
// V = double: the fp64 verification build, one env per thread.
// (The kernel is written for LANES envs per thread; a build with V = f2, TWO envs per thread in packed fp32x2 registers
// and one block per SM at 255 registers, was measured in round 2 — 106 vs 89 us per step, profiles/r2_step_kernel_pair_*
// — and dropped: with 7 warps per SM the kernel is bound by dependent-issue latency, not by issue slots. The packed
// instructions are used INSIDE an env instead: os2r_device.cuh, forward pass.)
// SJ / SC: the model's structure signature (os2r_device.cuh): the shipped URDFs' own, or the all-general one.
template <typename V, int N, int NC, int BLOCK, bool DAMPED, int MINB, uint32_t SJ, uint32_t SC>
__global__ void __launch_bounds__(BLOCK, MINB)
step_kernel(const __grid_constant__ ModelDev<typename VT<V>::S> M, const __grid_constant__ TaskDev K,
            StateDev<typename VT<V>::S> S, const __grid_constant__ StepIO IO, StatsDev *stats) {
    using T = typename VT<V>::S;
    constexpr int LANES = VT<V>::LANES;
    constexpr int EPB = BLOCK * LANES;             // envs per block
    const float *__restrict__ actions = IO.actions;
    float *__restrict__ obs = IO.obs;
    float *__restrict__ reward = IO.reward;
    uint8_t *__restrict__ done = IO.done;
    float *__restrict__ term_obs = IO.term_obs;
    int32_t *__restrict__ info = IO.info;
    static_assert(NC <= 6, "class byte: one bit per contact proxy, 2^NC + 1 sort keys");
    const int64_t NE = S.n_envs;
    using SL = ColdSlots<N, NC>;
    constexpr int ROWS = SL::ROWS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // Thread t steps the envs at sorted positions LANES*t.. of the block's window. Slots past the end of the batch (last
    // block only) sort last and shadow the last env instead of exiting, so that the block-wide barriers inside the
    // physics loop stay legal; they are skipped by the epilogue, before anything is written.
    const int64_t window = (int64_t)blockIdx.x * EPB;
    // Narrow blocks (two warps; batches of at most one wave of them) do not sort: such a batch is latency-bound — a step
    // lasts as long as the slowest warp — so regrouping lanes saves nothing there, while the class byte's DRAM round trip
    // and the sort's barriers sit in front of every other load of the launch.
    constexpr bool SORT = OS2R_SORT_NARROW || BLOCK > 64;
    int src[LANES];
    if constexpr (SORT) {
    int key[LANES];
#pragma unroll
    for (int h = 0; h < LANES; ++h) {
        const int64_t en = window + h * BLOCK + threadIdx.x;
        const bool nominal_valid = en < NE;
        key[h] = nominal_valid ? 1 + (int)(__ldcg(S.cls + en) & ((1u << NC) - 1u)) : 0;
        // While the class byte travels and the block sorts, pull the window's state lines towards the L2: which env a
        // thread will step is not known yet, but it is one of this window's, so the DRAM latency of the prologue loads
        // overlaps the sort instead of following it.
        if (nominal_valid) {
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
    }
    sorted_sources<BLOCK, LANES, (1 << NC) + 1>(key, src, reinterpret_cast<int *>(smem_raw));
    } else {
#pragma unroll
        for (int h = 0; h < LANES; ++h) src[h] = h * BLOCK + threadIdx.x;
    }
#ifdef OS2R_CHECKED
    {   // the sort must hand every window slot to exactly one thread
        __shared__ int owner[EPB];
#pragma unroll
        for (int h = 0; h < LANES; ++h) {
            OS2R_CHECK(src[h] >= 0 && src[h] < EPB, 0);
            if (src[h] >= 0 && src[h] < EPB) owner[src[h]] = LANES * threadIdx.x + h;
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < LANES; ++h)
            if (src[h] >= 0 && src[h] < EPB) OS2R_CHECK(owner[src[h]] == LANES * (int)threadIdx.x + h, 1);
        __syncthreads();
    }
#endif
    int64_t e[LANES];
    bool valid[LANES];
#pragma unroll
    for (int h = 0; h < LANES; ++h) {
        const int64_t e_raw = window + src[h];
        valid[h] = e_raw < NE;
        e[h] = valid[h] ? e_raw : NE - 1;
        OS2R_CHECK(e[h] >= 0 && e[h] < NE, 2);
    }
    const Cold<V, BLOCK> C{reinterpret_cast<V *>(smem_raw) + threadIdx.x};

    EnvRegs<V, N> E;
    // ---- prologue: ALL global loads are issued back to back (explicit ld.global, so the compiler may hoist
    //      them above the shared-memory stores that follow; generic loads were serialised LD -> STS -> LD ...,
    //      one DRAM round trip each) and only then scattered into registers / shared memory.
    V ld_qlo[N], ld_vlo[N], ld_mass[N], ld_damp[N], ld_fric[N], ld_lam[ROWS], ld_mu[NC];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        E.q_hi[i] = gather<V>(S.q_hi + i * NE, e);
        E.v[i] = gather<V>(S.qd + i * NE, e);
        ld_qlo[i] = gather<V>(S.q_lo + i * NE, e);
        ld_vlo[i] = gather<V>(S.qd_lo + i * NE, e);
        ld_mass[i] = gather<V>(S.mass_scale + i * NE, e);
        ld_damp[i] = gather<V>(S.damping + i * NE, e);
        ld_fric[i] = gather<V>(S.friction + i * NE, e);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) ld_lam[r] = gather<V>(S.lam + r * NE, e);
#pragma unroll
    for (int c = 0; c < NC; ++c) ld_mu[c] = gather<V>(S.mu + c * NE, e);
    E.gz = gather<V>(S.gravity_z, e);
    // episode bookkeeping and the previous action are only needed by the epilogue, but their loads are issued here with
    // all the others (one DRAM round trip for the lot) and the values parked in shared memory as bit patterns
    static_assert(LANES == 1, "the bookkeeping slots hold one env per thread");
    const T a_old0 = __ldcg(S.a_prev + e[0]), a_old1 = __ldcg(S.a_prev + NE + e[0]);
    const int steps_ld = __ldcg(S.steps + e[0]);
    const double ret_ld = __ldcg(S.ret + e[0]);
    float ax_[LANES], ay_[LANES], afin_[LANES];
#pragma unroll
    for (int h = 0; h < LANES; ++h) {
        float2 act = __ldcg(reinterpret_cast<const float2 *>(actions) + e[h]);
        // A non-finite action (the reference rejects it through `assert action_space.contains`, tasks/monopod.py:218)
        // applies no torque and takes the non-finite path of the epilogue: defined outputs, forced reset, counted.
        // fminf / fmaxf alone would silently turn a NaN into a full negative torque.
        const bool fin = isfinite(act.x) && isfinite(act.y);
        if (!fin) { act.x = 0.0f; act.y = 0.0f; }
        // ScenarIO clips force targets to +-max force (tasks/monopod.py:313-316)
        ax_[h] = fminf(1.0f, fmaxf(-1.0f, act.x));
        ay_[h] = fminf(1.0f, fmaxf(-1.0f, act.y));
        afin_[h] = fin ? 1.0f : 0.0f;
    }
    const V act_x = from_halves<V>(ax_), act_y = from_halves<V>(ay_);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        C(SL::QLO + i) = ld_qlo[i];
        C(SL::VLO + i) = ld_vlo[i];
        C(SL::MASS + i) = ld_mass[i];
        C(SL::DAMP + i) = ld_damp[i];
        C(SL::FRIC + i) = ld_fric[i] * M.dt;
        V tau = V(0);
        if (i == M.hip_dof) tau = act_x * M.max_torque[0];
        if (i == M.knee_dof) tau = act_y * M.max_torque[1];
        C(SL::TAU + i) = tau;
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) C(SL::LAM + r) = ld_lam[r];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        C(SL::MU + c) = ld_mu[c];
        C(SL::CX + 3 * c + 2) = V(0);   // defined even when substeps == 0 (reads as "near")
    }
    C(SL::ACT) = act_x;
    C(SL::ACT + 1) = act_y;
    C(SL::AFIN) = from_halves<V>(afin_);
    C(SL::AOLD) = a_old0;
    C(SL::AOLD + 1) = a_old1;
    C(SL::MISC) = int_as_real(T(0), steps_ld);
    C(SL::MISC + 1) = int_as_real(T(0), __double2loint(ret_ld));
    C(SL::MISC + 2) = int_as_real(T(0), __double2hiint(ret_ld));
#ifdef OS2R_CHECKED
    C(SL::COUNT) = V(12345.0f);          // guard word behind the last slot (the checked build allocates one more slot)
#endif

#pragma unroll 1
    for (int s = 0; s < M.substeps; ++s) {
        // Keep the block's warps in phase: all warps of an SM run the same ~40 KB loop body, and warps that drift
        // apart thrash the instruction caches.
        if (!(OS2R_SKIP_PHASE_BARRIERS & 1)) __syncthreads();
        physics_iteration<V, N, NC, DAMPED, SJ, SC, Cold<V, BLOCK>>(M, E, C);
    }
#ifdef OS2R_CHECKED
    OS2R_CHECK(half_of(C(SL::COUNT), 0) == (T)12345.0f, 4);
#endif
    // ---- epilogue (fp64, once per ENV): the hot registers are parked next to the cold slots and every env of the
    //      thread is finished from shared memory, one after the other, by the same (rolled) code
#pragma unroll
    for (int i = 0; i < N; ++i) { C(SL::QHI + i) = E.q_hi[i]; C(SL::VHI + i) = E.v[i]; }
    // what the warp-cooperative reset below needs from the per-env part (one env per thread: LANES == 1)
    bool do_reset = false, finished_ok = false;
    int64_t en_r = 0;
    int reset_idx_r = 0, cause_r = 0;
    double a_old_r[2] = {0.0, 0.0};
#pragma unroll 1
    for (int h = 0; h < LANES; ++h) {
        const bool ok = LANES == 1 ? valid[0] : (h ? valid[LANES - 1] : valid[0]);
        if (!ok) continue;
        const int64_t en = LANES == 1 ? e[0] : (h ? e[LANES - 1] : e[0]);
        const int steps_in = real_as_int(C.half(SL::MISC, h));
        int reset_idx = __ldcg(S.reset_id + en);
        const double ret_in = __hiloint2double(real_as_int(C.half(SL::MISC + 2, h)), real_as_int(C.half(SL::MISC + 1, h)));
        const double a_old[2] = {(double)C.half(SL::AOLD, h), (double)C.half(SL::AOLD + 1, h)};
        const float actx = (float)C.half(SL::ACT, h), acty = (float)C.half(SL::ACT + 1, h);
        double q[N], v[N];
        bool finite = C.half(SL::AFIN, h) != T(0);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            q[i] = (double)C.half(SL::QHI + i, h) + (double)C.half(SL::QLO + i, h);
            v[i] = (double)C.half(SL::VHI + i, h) + (double)C.half(SL::VLO + i, h);
            finite = finite && isfinite(q[i]) && isfinite(v[i]);
        }
        if (!finite) {   // nothing non-finite leaves the kernel: the step reports a zero state's observation, reward 0
#pragma unroll
            for (int i = 0; i < N; ++i) { q[i] = 0.0; v[i] = 0.0; }
        }
        const double a0[2] = {(double)actx, (double)acty};
        double o[OS2R_MAX_OBS];
        const bool task_done = observe<N, sizeof(T) == 4>(K, q, v, a_old, o);
        const double r = finite ? reward_fn(K.cfg, o, a0, a_old) : 0.0;
        const int D = K.cfg.obs_dim;
        int cause = (task_done && finite) ? 1 : 0;
        if (!finite) cause = 4;
        const int steps = steps_in + 1;
        const double ret = ret_in + r;
        if (K.cfg.max_episode_steps > 0 && steps >= K.cfg.max_episode_steps) cause |= 2;

        reward[en] = (float)r;
        done[en] = cause != 0;
        if (term_obs) for (int k = 0; k < D; ++k) term_obs[en * D + k] = (float)o[k];
        S.a_prev[en] = (T)actx;
        S.a_prev[NE + en] = (T)acty;

        if (cause && IO.term_count) {
            const int k = atomicAdd(IO.term_count, 1);
            OS2R_CHECK(k < IO.term_cap, 3);
            if (k < IO.term_cap) {
                int32_t *rec = IO.term_records + (int64_t)k * (D + 2);
                rec[0] = (int32_t)en;
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
            do_reset = true;     // done by the whole warp after this loop
        } else {
            // next step's sorting hint: which proxies ended within sort_margin of the ground
            unsigned cls = 0;
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (!(C.half(SL::CX + 3 * c + 2, h) - M.contact_radius[c] >= M.sort_margin)) cls |= 1u << c;
            S.cls[en] = (uint8_t)cls;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                S.q_hi[i * NE + en] = C.half(SL::QHI + i, h);
                S.q_lo[i * NE + en] = C.half(SL::QLO + i, h);
                S.qd[i * NE + en] = C.half(SL::VHI + i, h);
                S.qd_lo[i * NE + en] = C.half(SL::VLO + i, h);
            }
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr) S.lam[rr * NE + en] = C.half(SL::LAM + rr, h);
            S.steps[en] = steps;
            S.ret[en] = ret;
            for (int k = 0; k < D; ++k) obs[en * D + k] = (float)o[k];
        }
        finished_ok = true; en_r = en; reset_idx_r = reset_idx; cause_r = cause; a_old_r[0] = a_old[0]; a_old_r[1] = a_old[1];
    }
    // ---- in-kernel resets, warp-cooperative. A reset consumes ~28 uniforms = 21 Philox4x32-10 blocks; evaluated one
    //      after the other by the one lane whose env finished they were ~3 000 instructions at the very end of that
    //      warp's life, i.e. on the kernel's critical path whenever ANY env of the batch finishes (free_hip / HoppingV1:
    //      40 envs per step; measured 10 of 127 us per step). Here lane L evaluates block L for the finishing env and
    //      the owner collects the 42 uniforms by shuffle: same stream, same values, one Philox evaluation deep.
    {
        const unsigned lane = threadIdx.x & 31u;
        unsigned need = __ballot_sync(0xffffffffu, do_reset);
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1u;
            const uint64_t gid_mine = (uint64_t)(S.first_env_id + en_r);
            const uint32_t ep_mine = (do_reset && (int)lane == src) ? S.episode[en_r] + 1u : 0u;
            const uint32_t gid_lo = __shfl_sync(0xffffffffu, (uint32_t)gid_mine, src);
            const uint32_t gid_hi = __shfl_sync(0xffffffffu, (uint32_t)(gid_mine >> 32), src);
            const uint32_t ep = __shfl_sync(0xffffffffu, ep_mine, src);
            const uint64_t gid = ((uint64_t)gid_hi << 32) | gid_lo;
            double u0 = 0.0, u1 = 0.0;
            if (lane < OS2R_N_DRAW_BLOCKS) rng_uniform_pair(S.seed, gid, ep, lane, &u0, &u1);
            double table[2 * OS2R_N_DRAW_BLOCKS];
#pragma unroll
            for (int k = 0; k < OS2R_N_DRAW_BLOCKS; ++k) {
                table[2 * k] = __shfl_sync(0xffffffffu, u0, k);
                table[2 * k + 1] = __shfl_sync(0xffffffffu, u1, k);
            }
            if ((int)lane == src) {
                const DrawsTable draw{table};
                reset_idx_r = reset_env_global<T, N, NC>(K, S, en_r, a_old_r, obs + en_r * K.cfg.obs_dim, ep, draw);   // also marks the env "unknown" (all near)
            }
        }
    }
    if (finished_ok) {
        if (info) {
            info[2 * en_r] = reset_idx_r;
            info[2 * en_r + 1] = cause_r;
        }
        if (IO.reset_id8) IO.reset_id8[en_r] = (uint8_t)reset_idx_r;
    }
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
// launchers (dispatch on precision / build, n_dof, block width, damping)
// ------------------------------------------------------------------------------------------------
static inline int grid_for(int64_t n, int per_block) { return (int)((n + per_block - 1) / per_block); }

// (moving joints, contact proxies) combinations the kernels are instantiated for: the shipped models carry the three
// leg proxies (hip, knee, foot); the free-hip model `monopod` adds a fourth on the hip bracket (hip_link), which its free
// boom_connector joint can swing into the ground. (5, 3) stays available for models compiled without it.
#define OS2R_DISPATCH(n_dof, n_contacts, CALL)                                          \
    switch ((n_dof) * 16 + (n_contacts)) {                                              \
    case 2 * 16 + 3: { constexpr int N_ = 2, NC_ = 3; CALL; } break;                    \
    case 3 * 16 + 3: { constexpr int N_ = 3, NC_ = 3; CALL; } break;                    \
    case 4 * 16 + 3: { constexpr int N_ = 4, NC_ = 3; CALL; } break;                    \
    case 5 * 16 + 3: { constexpr int N_ = 5, NC_ = 3; CALL; } break;                    \
    case 5 * 16 + 4: { constexpr int N_ = 5, NC_ = 4; CALL; } break;                    \
    default: return cudaErrorInvalidValue;                                              \
    }
bool supported_shape(int n_dof, int n_contacts) {
    auto probe = [&]() -> cudaError_t { OS2R_DISPATCH(n_dof, n_contacts, (void)N_; (void)NC_); return cudaSuccess; };
    return probe() == cudaSuccess;
}

template <typename V, int N, int NC, int BLOCK>
static constexpr size_t step_smem_bytes() {
    constexpr int LANES = VT<V>::LANES;
    constexpr int NKEYS = (1 << NC) + 1;
#ifdef OS2R_CHECKED
    size_t cold = (size_t)(ColdSlots<N, NC>::COUNT + 1) * BLOCK * sizeof(V);     // + the guard slot
#else
    size_t cold = (size_t)ColdSlots<N, NC>::COUNT * BLOCK * sizeof(V);
#endif
    size_t sort = (size_t)(((NKEYS * (BLOCK / 32) * LANES + 31) / 32) * 32 + BLOCK * LANES) * sizeof(int);
    return cold > sort ? cold : sort;
}

// Threads per block of the step kernel for a batch. Wide blocks (7 warps, two per SM at 128 registers) give the lane
// sort enough envs to fill whole warps with one contact class. Batches that fit ONE wave of the narrow build (four 2-warp
// blocks per SM at up to 255 registers: 37 888 envs on 148 SMs) run it instead — more registers, more instruction-level
// parallelism, and no SM that holds two wide blocks while its neighbour holds one (measured, fixed_hip steady state:
// 36 864 envs 67.0 us narrow / 84.3 wide; 40 960 envs 104.7 narrow (two waves) / 85.5 wide). fp64: narrow blocks only.
int step_block_threads(int build, int64_t n_envs, int sm_count) {
    if (build == OS2R_BUILD_F32 && n_envs > (int64_t)sm_count * 4 * OS2R_BLOCK) return OS2R_BLOCK_WIDE;
    return OS2R_BLOCK;
}

namespace {

struct StepFn {          // one instantiation of the step kernel
    const void *fn;
    size_t smem;
    int envs_per_block;
};

// Structure signatures of the shipped models (gym_os2r_b200/models/assets/*.urdf), as model_signature() computes
// them: monopod-simple (2 joints), monopod-fixed (3), monopod-fixed_hip (4), monopod (5, with and without the bracket
// proxy), and whether that model's joints are damped. tests/test_capi_cpu.py::test_shipped_models_run_the_specialised_kernels
// keeps the table honest; any other model runs the all-general instantiation.
template <int N, int NC> struct Shipped { static constexpr uint32_t J = 0xffffffffu, C = 0xffffffffu; static constexpr bool DAMPED = false; };   // none
template <> struct Shipped<2, 3> { static constexpr uint32_t J = OS2R_SHIPPED_J23, C = OS2R_SHIPPED_C23; static constexpr bool DAMPED = true; };
template <> struct Shipped<3, 3> { static constexpr uint32_t J = OS2R_SHIPPED_J33, C = OS2R_SHIPPED_C33; static constexpr bool DAMPED = true; };
template <> struct Shipped<4, 3> { static constexpr uint32_t J = OS2R_SHIPPED_J43, C = OS2R_SHIPPED_C43; static constexpr bool DAMPED = false; };
template <> struct Shipped<5, 3> { static constexpr uint32_t J = OS2R_SHIPPED_J53, C = OS2R_SHIPPED_C53; static constexpr bool DAMPED = true; };
template <> struct Shipped<5, 4> { static constexpr uint32_t J = OS2R_SHIPPED_J54, C = OS2R_SHIPPED_C54; static constexpr bool DAMPED = true; };

template <typename V, int N, int NC, int BLOCK, bool DAMPED, bool SPECIAL>
StepFn step_fn() {
    // two resident 7-warp blocks per SM for the float build (128 registers per thread); everything else: no occupancy target
    constexpr int MINB = (VT<V>::LANES == 1 && sizeof(typename VT<V>::S) == 4 && BLOCK == OS2R_BLOCK_WIDE) ? OS2R_WIDE_MINB : 1;
    constexpr uint32_t SJ = SPECIAL ? Shipped<N, NC>::J : generic_joint_signature(N);
    constexpr uint32_t SC = SPECIAL ? Shipped<N, NC>::C : generic_contact_signature(NC);
    return StepFn{(const void *)step_kernel<V, N, NC, BLOCK, DAMPED, MINB, SJ, SC>, step_smem_bytes<V, N, NC, BLOCK>(), BLOCK * VT<V>::LANES};
}
template <typename V, int N, int NC, int BLOCK>
StepFn step_fn_d(bool damped, uint32_t sj, uint32_t sc) {
    // the specialised fp32 kernels exist for the damping variant the shipped model of this shape runs (Shipped::DAMPED:
    // monopod-fixed_hip carries no joint damping, the other URDFs do); the fp64 verification build keeps one (damped)
    // instantiation per signature — with zero damping its second factor equals the first
    const bool special = sj == Shipped<N, NC>::J && sc == Shipped<N, NC>::C;
    if constexpr (sizeof(typename VT<V>::S) == 8) {
        return special ? step_fn<V, N, NC, BLOCK, true, true>() : step_fn<V, N, NC, BLOCK, true, false>();
    } else {
        constexpr bool SD = Shipped<N, NC>::DAMPED;
        if (special && damped == SD) return step_fn<V, N, NC, BLOCK, SD, true>();
        return damped ? step_fn<V, N, NC, BLOCK, true, false>() : step_fn<V, N, NC, BLOCK, false, false>();
    }
}
template <typename V>
cudaError_t pick(int n_dof, int n_contacts, int block, bool damped, uint32_t sj, uint32_t sc, StepFn *out) {
    if (block == OS2R_BLOCK_WIDE) {
        if constexpr (sizeof(typename VT<V>::S) == 4) {
            OS2R_DISPATCH(n_dof, n_contacts, *out = (step_fn_d<V, N_, NC_, OS2R_BLOCK_WIDE>(damped, sj, sc)));
            return cudaSuccess;
        }
        return cudaErrorInvalidValue;
    }
    if (block != OS2R_BLOCK) return cudaErrorInvalidValue;
    OS2R_DISPATCH(n_dof, n_contacts, *out = (step_fn_d<V, N_, NC_, OS2R_BLOCK>(damped, sj, sc)));
    return cudaSuccess;
}
cudaError_t pick_build(int build, int n_dof, int n_contacts, int block, bool damped, uint32_t sj, uint32_t sc, StepFn *out) {
    switch (build) {
    case OS2R_BUILD_F32: return pick<float>(n_dof, n_contacts, block, damped, sj, sc, out);
    case OS2R_BUILD_F64: return pick<double>(n_dof, n_contacts, block, true, sj, sc, out);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace

bool shipped_signature(int n_dof, int n_contacts, uint32_t *sj, uint32_t *sc) {
    auto probe = [&]() -> cudaError_t { OS2R_DISPATCH(n_dof, n_contacts, (*sj = Shipped<N_, NC_>::J, *sc = Shipped<N_, NC_>::C)); return cudaSuccess; };
    return probe() == cudaSuccess;
}

template <typename T>
cudaError_t launch_step(int build, int n_dof, int n_contacts, int block, uint32_t sj, uint32_t sc, const ModelDev<T> &M,
                        const TaskDev &K, const StateDev<T> &S, const StepIO &io, StatsDev *stats, cudaStream_t stream) {
    if ((build == OS2R_BUILD_F64) != (sizeof(T) == 8)) return cudaErrorInvalidValue;
    StepFn f;
    cudaError_t e = pick_build(build, n_dof, n_contacts, block, M.any_damping != 0, sj, sc, &f);
    if (e != cudaSuccess) return e;
    // blocks that need more than 48 KB of dynamic shared memory were opted in by prepare_step (once per handle,
    // on the handle's device: the attribute is per device, a process can hold handles on several GPUs)
    void *args[] = {(void *)&M, (void *)&K, (void *)&S, (void *)&io, (void *)&stats};
    return cudaLaunchKernel(f.fn, dim3(grid_for(S.n_envs, f.envs_per_block)), dim3(block), args, f.smem, stream);
}

template <typename T>
cudaError_t launch_reset(int n_dof, int n_contacts, const TaskDev &K, const StateDev<T> &S, const uint8_t *mask,
                         float *obs, cudaStream_t stream) {
    OS2R_DISPATCH(n_dof, n_contacts, (reset_kernel<T, N_, NC_><<<grid_for(S.n_envs, OS2R_BLOCK), OS2R_BLOCK, 0, stream>>>(K, S, mask, obs)));
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_init(const TaskDev &K, const StateDev<T> &S, double nominal_gz, cudaStream_t stream) {
    init_kernel<T><<<grid_for(S.n_envs, OS2R_BLOCK), OS2R_BLOCK, 0, stream>>>(K, S, nominal_gz);
    return cudaGetLastError();
}

// Opt the step kernels this handle can launch (damped and undamped build) into their dynamic shared memory size on the
// CURRENT device. Called by os2r_create under its device guard.
cudaError_t prepare_step(int build, int n_dof, int n_contacts, int block, uint32_t sj, uint32_t sc) {
    for (int damped = 0; damped < 2; ++damped) {
        StepFn f;
        cudaError_t e = pick_build(build, n_dof, n_contacts, block, damped != 0, sj, sc, &f);
        if (e != cudaSuccess) return e;
        if (f.smem > 48 * 1024) {
            e = cudaFuncSetAttribute(f.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f.smem);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

cudaError_t step_kernel_attributes(int build, int n_dof, int n_contacts, int block, bool damped, uint32_t sj, uint32_t sc,
                                   cudaFuncAttributes *attr, int *blocks_per_sm, int *envs_per_block) {
    StepFn f;
    cudaError_t e = pick_build(build, n_dof, n_contacts, block, damped, sj, sc, &f);
    if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(attr, f.fn);
    if (e != cudaSuccess) return e;
    if (envs_per_block) *envs_per_block = f.envs_per_block;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, f.fn, block, f.smem);
}

cudaError_t read_check_counters(unsigned long long out[8], bool clear) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_check, 8 * sizeof(unsigned long long));
    if (e != cudaSuccess) return e;
#ifdef OS2R_CHECKED
    out[7] = 1;
#else
    out[7] = 0;
#endif
    if (clear) {
        const unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        e = cudaMemcpyToSymbol(g_check, zero, sizeof(zero));
    }
    return e;
}

cudaError_t launch_fma_peak(float *out, int blocks, int iters, cudaStream_t stream) {
    fma_peak_kernel<<<blocks, 256, 0, stream>>>(out, iters, 1.0f);
    return cudaGetLastError();
}

#define OS2R_INSTANTIATE(T)                                                                                   \
    template cudaError_t launch_step<T>(int, int, int, int, uint32_t, uint32_t, const ModelDev<T> &, const TaskDev &, const StateDev<T> &, \
                                        const StepIO &, StatsDev *, cudaStream_t);                            \
    template cudaError_t launch_reset<T>(int, int, const TaskDev &, const StateDev<T> &, const uint8_t *,     \
                                         float *, cudaStream_t);                                              \
    template cudaError_t launch_init<T>(const TaskDev &, const StateDev<T> &, double, cudaStream_t);
OS2R_INSTANTIATE(float)
OS2R_INSTANTIATE(double)

}  // namespace os2r
