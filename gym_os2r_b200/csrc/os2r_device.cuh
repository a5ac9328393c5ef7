// os2r_device.cuh — device-side monopod physics + task epilogue for sm_100a.
//
// One thread steps one environment through `substeps` physics iterations and the fused task
// epilogue (observation, reward, termination, auto-reset, randomiser draws). All model constants
// arrive as a __grid_constant__ kernel parameter (constant bank, broadcast to the warp); per-env
// state lives in HBM as structure-of-arrays [field][env] so every load/store is coalesced.
//
// Formulation (deliberately different from the CPU oracle's body-frame ABA):
//   * every joint is normalised on the host to rotate about its child-frame x axis;
//   * ONE forward pass over the chain in world axes computes kinematics, classical Newton-Euler
//     velocities / accelerations (qdd = 0), and accumulates the joint-space mass matrix and bias
//     directly:  M = sum_b [m Jv^T Jv + Jw^T Iw Jw],  h = sum_b [Jv^T m a_c + Jw^T (Iw al + w x Iw w)].
//     Nothing per-body survives the pass (register footprint), and every term is a difference of nearby
//     points or a positive contribution (no reference point, no fp32 m|r|^2 cancellation);
//   * qdd from a Cholesky solve of (M + dt*D) (implicit joint damping, as DART);
//   * constraints (Coulomb joint friction rows + per-contact normal / 2 friction rows) solved by
//     projected Gauss-Seidel in Cholesky-whitened velocity coordinates z = L^T v: each row needs a
//     single n-vector G_r = L^-1 J_r^T (row velocity = G_r.z, update z += G_r*dlambda,
//     A_rr = |G_r|^2), which halves the register footprint versus storing J_r and M^-1 J_r^T;
//   * positions AND velocities are integrated with compensated (hi, lo) float pairs in the fp32 build
//     (velocity rounding, amplified by the dynamics, was the dominant fp32 drift: measured 10x).
//
// Reference semantics restated: gym_os2r/runtimes/gazebo_runtime.py:65-97 (10x zero-order hold),
// gym_os2r/tasks/monopod.py:202-298 (torque map, observation, done), gym_os2r/rewards/,
// gym_os2r/randomizers/monopod.py:56-135,182-215, gym_os2r/utils/reset.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/os2r.h"

// A/B switch for the phase barriers of the physics loop (bit 0: loop top, bit 1: after the forward pass, bit 2: in front
// of the contact rows) for tools/kprobe.py with a variant library (OS2R_LIB). The product build keeps all three.
#ifndef OS2R_SKIP_PHASE_BARRIERS
#define OS2R_SKIP_PHASE_BARRIERS 0
#endif

namespace os2r {

// ------------------------------------------------------------------------------------------------
// device-side constant tables
// ------------------------------------------------------------------------------------------------
template <typename T>
struct ModelDev {
    T tree_R[OS2R_MAX_DOF][9];   // row-major, joint axis already normalised to child x
    T tree_p[OS2R_MAX_DOF][3];
    T mass[OS2R_MAX_DOF];
    T com[OS2R_MAX_DOF][3];
    T inertia[OS2R_MAX_DOF][6];  // xx yy zz xy xz yz
    T contact_pos[OS2R_MAX_CONTACTS][3];
    T contact_radius[OS2R_MAX_CONTACTS];
    T dt, erp_over_dt, max_erv, cfm_contact, cfm_joint;
    T cfm1_contact, cfm1_joint;  // 1 + cfm
    T kc1, kj1;                  // 1 / (1 + cfm)  (host-side: an IEEE division per physics iteration otherwise)
    T root_mass_term, root_inertia_term;  // see root_spin below
    T pgs_tol2;                  // squared energy-norm tolerance of the sweeps (os2r_model.pgs_tol)
    T sort_margin;               // ground clearance below which a contact proxy counts as "near" (lane sorting hint)
    T max_torque[2];
    int32_t contact_body[OS2R_MAX_CONTACTS];
    int32_t hip_dof, knee_dof;
    int32_t substeps, pgs_iters;
    int32_t pgs_joint_sweeps;    // sweeps of an iteration that include the joint-friction rows (0: all)
    int32_t any_damping;         // 0 when every joint's nominal damping is 0 (skip 2nd factorisation)
    int32_t root_spin;           // 1: body 0 hangs off the world and turns about an axis parallel to gravity (the yaw
                                 //    pivot). Its whole contribution to the dynamics is then a CONSTANT added to M[0][0]
                                 //    (mass * |a x d|^2 * mass_scale + a.I.a: rotation about a leaves both unchanged) and
                                 //    nothing to the bias (gravity || a, centrifugal force radial): the host folds it.
};

struct TaskDev {                 // epilogue + reset configuration (fp64: evaluated once per env step)
    os2r_task_cfg cfg;
    double nominal_damping[OS2R_MAX_DOF];
    double nominal_friction[OS2R_MAX_DOF];
    double nominal_mu[OS2R_MAX_CONTACTS];
    double obs_scale[OS2R_MAX_OBS];   // 2 / (obs_high - obs_low): the fp32 build normalises with one FMA instead of a division
    int32_t role_dof[OS2R_N_ROLES];
    int32_t n_dof, n_contacts;
};

// SoA views of the per-env state in HBM. Every array is [count][n_envs].
template <typename T>
struct StateDev {
    T *q_hi, *q_lo, *qd, *qd_lo; // [n_dof][N]  (lo = compensation terms of the fp32 build)
    T *lam;                      // [rows][N]   warm-start impulses
    T *a_prev;                   // [2][N]      last applied action
    T *mass_scale, *damping, *friction;  // [n_dof][N]
    T *mu;                       // [n_contacts][N]
    T *gravity_z;                // [N]
    int32_t *steps;              // [N] steps in the current episode
    uint32_t *episode;           // [N] episode counter (RNG counter word)
    int32_t *reset_id;           // [N]
    double *ret;                 // [N] return of the current episode
    uint8_t *cls;                // [N] bit c: contact proxy c was near the ground after the last step. A pure
                                 //     scheduling hint (which lanes share a warp); never changes a result.
    int64_t n_envs;
    int64_t first_env_id;
    uint64_t seed;
};

// I/O buffers of one step launch. actions/obs/reward/done are mandatory; the rest is optional (null = not wanted).
struct StepIO {
    const float *actions;        // [N,2]
    float *obs;                  // [N,D]
    float *reward;               // [N]
    uint8_t *done;               // [N] 0/1
    float *term_obs;             // [N,D] dense terminal observation (== obs for envs that did not finish)
    int32_t *info;               // [N,2] {reset id after the step, done cause bits}
    // compact outputs for the host path (os2r_step_host_packed): one byte of reset id per env, and one record
    // {env index, cause, terminal observation[D]} per FINISHED env appended through an atomic counter
    uint8_t *reset_id8;          // [N]
    int32_t *term_count;         // [1], zeroed by the caller before the launch
    int32_t *term_records;       // [term_cap][D + 2] words
    int32_t term_cap;
};

struct StatsDev {                // device-side accumulators (os2r_stats without env_steps)
    unsigned long long episodes, done_task, done_timelimit, nonfinite_resets;
    double sum_return, sum_length;
};

// ------------------------------------------------------------------------------------------------
// value types: the physics is written once over a value type V
//   float / double : one env per thread (double = the verification build)
//   f2             : TWO envs per thread in one 64-bit register pair, arithmetic through the packed fp32x2
//                    instructions of sm_100 (PTX fma/add/sub/mul.rn.f32x2 -> SASS FFMA2 / FADD2 / FMUL2). One FFMA2
//                    issues in ONE scheduler slot and does the work of two FFMAs (measured on B200,
//                    tools/microbench/ffma2_probe.cu: same 4.6-cycle dependent latency as FFMA, two FMA-pipe cycles per
//                    warp instruction, i.e. the same peak flops through HALF the issue slots). Scalars (model constants
//                    in the constant bank, immediates) broadcast to both halves inside the instruction (`R.F32` /
//                    `UR.F32` operand forms): no duplication needed. MEASURED AND NOT SHIPPED: with one 7-warp block per
//                    SM at 255 registers the pair build ran 106 us per step against 89 for one env per thread
//                    (the kernel is bound by dependent-issue latency once 1.75 warps share a
//                    scheduler); no kernel is instantiated on f2 any more, the type stays for the record and for A/B.
// Per-half operations without a packed instruction (min / max / compare / select / MUFU) run once per half.
// ------------------------------------------------------------------------------------------------
struct f2 {
    unsigned long long r;
    __device__ __forceinline__ f2() {}
    __device__ __forceinline__ f2(float s) { asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(s)); }
    __device__ __forceinline__ f2(float lo, float hi) { asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); }
    __device__ __forceinline__ float lo() const { float a; asm("{ .reg .f32 t; mov.b64 {%0, t}, %1; }" : "=f"(a) : "l"(r)); return a; }
    __device__ __forceinline__ float hi() const { float b; asm("{ .reg .f32 t; mov.b64 {t, %0}, %1; }" : "=f"(b) : "l"(r)); return b; }
};
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.r) : "l"(a.r), "l"(b.r)); return d; }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { f2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.r) : "l"(a.r), "l"(b.r)); return d; }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.r) : "l"(a.r), "l"(b.r)); return d; }
__device__ __forceinline__ f2 operator-(f2 a) { return f2(-a.lo(), -a.hi()); }   // folds into the consumer's operand modifier
__device__ __forceinline__ f2 &operator+=(f2 &a, f2 b) { a = a + b; return a; }
__device__ __forceinline__ f2 &operator-=(f2 &a, f2 b) { a = a - b; return a; }
__device__ __forceinline__ f2 &operator*=(f2 &a, f2 b) { a = a * b; return a; }

template <typename V> struct VT;             // traits: scalar type, envs per thread, mask type
template <> struct VT<float>  { using S = float;  using M = bool; static constexpr int LANES = 1; };
template <> struct VT<double> { using S = double; using M = bool; static constexpr int LANES = 1; };
struct m2 { bool a, b; };
template <> struct VT<f2>     { using S = float;  using M = m2;   static constexpr int LANES = 2; };

// half h of a value (h is a compile-time constant at every call site)
__device__ __forceinline__ float half_of(float v, int) { return v; }
__device__ __forceinline__ double half_of(double v, int) { return v; }
__device__ __forceinline__ float half_of(f2 v, int h) { return h ? v.hi() : v.lo(); }

__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ f2 fma_t(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.r) : "l"(a.r), "l"(b.r), "l"(c.r)); return d; }
__device__ __forceinline__ float fmin_t(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double fmin_t(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ f2 fmin_t(f2 a, f2 b) { return f2(fminf(a.lo(), b.lo()), fminf(a.hi(), b.hi())); }
__device__ __forceinline__ float fmax_t(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double fmax_t(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ f2 fmax_t(f2 a, f2 b) { return f2(fmaxf(a.lo(), b.lo()), fmaxf(a.hi(), b.hi())); }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
// 1/sqrt(x) to ~1 ulp: MUFU.RSQ + one Newton step (shorter dependency chain than sqrtf followed by a division)
__device__ __forceinline__ float rsqrt_t(float x) {
    const float r = rsqrtf(x);
    return fmaf(r, fmaf(-0.5f * x * r, r, 0.5f), r);
}
__device__ __forceinline__ double rsqrt_t(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ f2 rsqrt_t(f2 x) {   // same operations per half as the scalar routine, Newton step packed
    const f2 r(rsqrtf(x.lo()), rsqrtf(x.hi()));
    return fma_t(r, fma_t(f2(-0.5f) * x * r, r, f2(0.5f)), r);
}
// MUFU.RCP (~1 ulp), no IEEE fix-up / denormal slow path: only used for the row relaxation factors 1/(A(1+cfm)),
// whose rounding moves the sweep's fixed point by cfm * ulp (the fixed point itself does not depend on the factor)
__device__ __forceinline__ float rcp_t(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ double rcp_t(double x) { return 1.0 / x; }
__device__ __forceinline__ f2 rcp_t(f2 x) { return f2(rcp_t(x.lo()), rcp_t(x.hi())); }

// masks: per-env predicates (a pair build carries one per half)
__device__ __forceinline__ bool gt_t(float a, float b) { return a > b; }
__device__ __forceinline__ bool gt_t(double a, double b) { return a > b; }
__device__ __forceinline__ m2 gt_t(f2 a, f2 b) { return m2{a.lo() > b.lo(), a.hi() > b.hi()}; }
__device__ __forceinline__ bool le_t(float a, float b) { return a <= b; }
__device__ __forceinline__ bool le_t(double a, double b) { return a <= b; }
__device__ __forceinline__ m2 le_t(f2 a, f2 b) { return m2{a.lo() <= b.lo(), a.hi() <= b.hi()}; }
__device__ __forceinline__ bool any_t(bool m) { return m; }
__device__ __forceinline__ bool any_t(m2 m) { return m.a || m.b; }
__device__ __forceinline__ bool and_not(bool m, bool k) { return m && !k; }
__device__ __forceinline__ m2 and_not(m2 m, m2 k) { return m2{m.a && !k.a, m.b && !k.b}; }
__device__ __forceinline__ float sel_t(bool m, float a, float b) { return m ? a : b; }
__device__ __forceinline__ double sel_t(bool m, double a, double b) { return m ? a : b; }
__device__ __forceinline__ f2 sel_t(m2 m, f2 a, f2 b) { return f2(m.a ? a.lo() : b.lo(), m.b ? a.hi() : b.hi()); }
template <typename V> __device__ __forceinline__ typename VT<V>::M all_true();
template <> __device__ __forceinline__ bool all_true<float>() { return true; }
template <> __device__ __forceinline__ bool all_true<double>() { return true; }
template <> __device__ __forceinline__ m2 all_true<f2>() { return m2{true, true}; }

// sin / cos of a compensated angle hi + lo: sin(hi+lo) = s + lo*c, cos(hi+lo) = c - lo*s.
template <typename T>
struct SinCos { T s, c; };       // returned by value: stays in registers across the call (pointers would go via the stack)
__device__ __forceinline__ void sincos_t(float x, float *s, float *c) { sincosf(x, s, c); }
__device__ __forceinline__ void sincos_t(double x, double *s, double *c) { sincos(x, s, c); }
// Library routine, deliberately NOT inlined (its range-reduction slow path is ~150 instructions): the fp32 builds only
// reach it beyond 1e5 rad, the fp64 build calls it once per joint.
template <typename T>
__device__ __noinline__ SinCos<T> joint_sincos(T hi, T lo) {
    T s, c;
    sincos_t(hi, &s, &c);
    return SinCos<T>{s + lo * c, c - lo * s};
}
// fp32 sin/cos for the forward pass, inlined and evaluated for all joints of an env side by side (independent
// chains the compiler interleaves; an out-of-line sincosf call per joint was 7 % of the instructions but 13 % of the
// stall samples of the contact-free step). Three-term Cody-Waite reduction by pi/2 (the same scheme as the library's
// fast path, exact enough for |x| < 1e5; beyond that the library routine is called) and the cephes minimax
// polynomials on [-pi/4, pi/4]: <= 1 ulp. Written over V so that the pair build runs reduction and polynomials packed;
// only the rounding to the quadrant and the quadrant selection run per half.
__device__ __forceinline__ float rint_t(float x) { return rintf(x); }
__device__ __forceinline__ f2 rint_t(f2 x) { return f2(rintf(x.lo()), rintf(x.hi())); }
__device__ __forceinline__ void quadrant_fix(float j, float sr, float cr, float *sn, float *cs) {
    const int q = (int)j;
    const float s0 = (q & 1) ? cr : sr, c0 = (q & 1) ? sr : cr;
    *sn = (q & 2) ? -s0 : s0;
    *cs = ((q + 1) & 2) ? -c0 : c0;
}
__device__ __forceinline__ void quadrant_fix(f2 j, f2 sr, f2 cr, f2 *sn, f2 *cs) {
    float s0, c0, s1, c1;
    quadrant_fix(j.lo(), sr.lo(), cr.lo(), &s0, &c0);
    quadrant_fix(j.hi(), sr.hi(), cr.hi(), &s1, &c1);
    *sn = f2(s0, s1);
    *cs = f2(c0, c1);
}
template <typename V>
__device__ __forceinline__ void sincos_fast(V x, V *sn, V *cs) {
    const V j = rint_t(x * V(0.636619772367581343f));          // x * 2/pi
    V r = fma_t(j, V(-1.57079601287841796875f), x);            // pi/2 split in three parts
    r = fma_t(j, V(-3.1391647326017846353352069854736328125e-7f), r);
    r = fma_t(j, V(-5.390302529957764765544681040410068817436695098876953125e-15f), r);
    const V r2 = r * r;
    V ps = fma_t(r2, V(-1.9515295891e-4f), V(8.3321608736e-3f));
    ps = fma_t(ps, r2, V(-1.6666654611e-1f));
    const V sr = fma_t(ps * r2, r, r);                         // sin r
    V pc = fma_t(r2, V(2.443315711809948e-5f), V(-1.388731625493765e-3f));
    pc = fma_t(pc, r2, V(4.166664568298827e-2f));
    const V cr = fma_t(pc * r2, r2, fma_t(r2, V(-0.5f), V(1.0f))); // cos r
    quadrant_fix(j, sr, cr, sn, cs);
}
// all joint angles of the thread's env(s): sin / cos of hi + lo
__device__ __forceinline__ bool beyond_fast_range(float q) { return !(fabsf(q) < 1.0e5f); }
__device__ __forceinline__ bool beyond_fast_range(f2 q) { return !(fabsf(q.lo()) < 1.0e5f) || !(fabsf(q.hi()) < 1.0e5f); }
__device__ __forceinline__ SinCos<float> slow_sincos(float hi, float lo) { return joint_sincos<float>(hi, lo); }
__device__ __forceinline__ SinCos<f2> slow_sincos(f2 hi, f2 lo) {
    const SinCos<float> a = joint_sincos<float>(hi.lo(), lo.lo()), b = joint_sincos<float>(hi.hi(), lo.hi());
    return SinCos<f2>{f2(a.s, b.s), f2(a.c, b.c)};
}

// bit-preserving int <-> float for parking integers in the shared-memory slots
__device__ __forceinline__ float int_as_real(float, int v) { return __int_as_float(v); }
__device__ __forceinline__ double int_as_real(double, int v) { return __hiloint2double(0, v); }
__device__ __forceinline__ int real_as_int(float v) { return __float_as_int(v); }
__device__ __forceinline__ int real_as_int(double v) { return __double2loint(v); }

#define OS2R_CROSS(o, a, b)                    \
    do {                                       \
        (o)[0] = (a)[1] * (b)[2] - (a)[2] * (b)[1]; \
        (o)[1] = (a)[2] * (b)[0] - (a)[0] * (b)[2]; \
        (o)[2] = (a)[0] * (b)[1] - (a)[1] * (b)[0]; \
    } while (0)
// o += a x b, accumulated term by term: two FMAs per component. Written as o += (a1*b2 - a2*b1) the compiler has to
// respect the parentheses (FMUL + FFMA + FADD); the forward pass is issue-bound, so the saved instruction counts.
#define OS2R_CROSS_ACC(o, a, b)                                               \
    do {                                                                      \
        (o)[0] = fma_t((a)[1], (b)[2], (o)[0]); (o)[0] = fma_t(-(a)[2], (b)[1], (o)[0]); \
        (o)[1] = fma_t((a)[2], (b)[0], (o)[1]); (o)[1] = fma_t(-(a)[0], (b)[2], (o)[1]); \
        (o)[2] = fma_t((a)[0], (b)[1], (o)[2]); (o)[2] = fma_t(-(a)[1], (b)[0], (o)[2]); \
    } while (0)
// acc += a . b as a chain of three FMAs (instead of FMUL + 2 FFMA + FADD)
#define OS2R_DOT_ACC(acc, a, b)                   \
    do {                                          \
        (acc) = fma_t((a)[0], (b)[0], (acc));     \
        (acc) = fma_t((a)[1], (b)[1], (acc));     \
        (acc) = fma_t((a)[2], (b)[2], (acc));     \
    } while (0)
#define OS2R_DOT(a, b) ((a)[0] * (b)[0] + (a)[1] * (b)[1] + (a)[2] * (b)[2])

// symmetric 3x3 (xx yy zz xy xz yz) times vector
#define OS2R_SYMV(o, S, v)                                        \
    do {                                                          \
        (o)[0] = (S)[0] * (v)[0] + (S)[3] * (v)[1] + (S)[4] * (v)[2]; \
        (o)[1] = (S)[3] * (v)[0] + (S)[1] * (v)[1] + (S)[5] * (v)[2]; \
        (o)[2] = (S)[4] * (v)[0] + (S)[5] * (v)[1] + (S)[2] * (v)[2]; \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Philox4x32-10, identical stream layout to the oracle (key = seed, counter = env, episode, block)
// ------------------------------------------------------------------------------------------------
__device__ inline void philox4x32(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t block, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)env_id, c1 = (uint32_t)(env_id >> 32), c2 = episode, c3 = block;
#pragma unroll 1
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ inline double rng_uniform(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t k) {
    uint32_t x[4];
    philox4x32(seed, env_id, episode, k >> 1, x);
    uint32_t a = (k & 1) ? x[2] : x[0], b = (k & 1) ? x[3] : x[1];
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
// both uniforms of one Philox block: draws 2*block and 2*block + 1
__device__ inline void rng_uniform_pair(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t block, double *u0, double *u1) {
    uint32_t x[4];
    philox4x32(seed, env_id, episode, block, x);
    *u0 = ((double)(x[0] >> 5) * 67108864.0 + (double)(x[1] >> 6)) * (1.0 / 9007199254740992.0);
    *u1 = ((double)(x[2] >> 5) * 67108864.0 + (double)(x[3] >> 6)) * (1.0 / 9007199254740992.0);
}
__device__ inline void rng_normal2(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t k, double z[2]) {
    double u1 = rng_uniform(seed, env_id, episode, k), u2 = rng_uniform(seed, env_id, episode, k + 1);
    double r = sqrt(-2.0 * log(1.0 - u1));
    double s, c;
    sincos(6.283185307179586476925 * u2, &s, &c);
    z[0] = r * c; z[1] = r * s;
}
// Where the uniforms of a reset come from: straight from Philox (one block per call: the thread-per-env reset kernel), or
// from a table the lanes of a warp filled together (the step kernel's in-kernel reset: lane L evaluates block L, so the
// warp runs ONE Philox evaluation instead of the ~28 a reset consumes one after the other on its critical path).
#define OS2R_N_DRAW_BLOCKS 21    /* draws 0 .. 41 (DRAW_GRAVITY + 1), two per Philox block */
struct DrawsPhilox {
    uint64_t seed, gid;
    uint32_t ep;
    __device__ double operator()(uint32_t k) const { return rng_uniform(seed, gid, ep, k); }
};
struct DrawsTable {
    const double *u;             // [2 * OS2R_N_DRAW_BLOCKS]
    __device__ double operator()(uint32_t k) const { return u[k]; }
};
template <typename D>
__device__ inline void draw_normal2(const D &draw, uint32_t k, double z[2]) {   // same arithmetic as rng_normal2
    double u1 = draw(k), u2 = draw(k + 1);
    double r = sqrt(-2.0 * log(1.0 - u1));
    double s, c;
    sincos(6.283185307179586476925 * u2, &s, &c);
    z[0] = r * c; z[1] = r * s;
}
enum { DRAW_RESET = 0, DRAW_PITCH = 1, DRAW_NOISE = 2, DRAW_LAYSIDE = 4, DRAW_DIR = 5, DRAW_YAW = 6,
       DRAW_SIMPLE_HIP = 7, DRAW_SIMPLE_KNEE = 8, DRAW_PARAMS = 10 /* .. 10 + 3*MAX_DOF + MAX_CONTACTS */,
       DRAW_GRAVITY = 40 /*,41*/ };
#define OS2R_EPISODE_GRAVITY 0xFFFFFFFFu

// ------------------------------------------------------------------------------------------------
// one physics iteration
// ------------------------------------------------------------------------------------------------
// Per-thread data that is touched only a few times per physics iteration lives in shared memory,
// laid out [slot][thread] (bank = thread, conflict-free): warm-start impulses, randomised parameters,
// the compensation terms of the (hi, lo) state pairs, torques and the contact-sphere centres.
// Only what every phase needs (q_hi, v, gravity) stays in registers. A slot holds one V: a float / double, or in the
// pair build the f2 of the thread's two envs (8-byte accesses, half h of slot k is the float at byte offset 4h).
// Contact proxies whose constraint rows (G, 1/A, b) live in SHARED MEMORY between their set-up and the sweeps instead
// of registers (bit c = proxy c). The sweep loop of the 5-DoF model with all four proxies in registers needs ~130 live
// registers against the 128 two resident 7-warp blocks allow (640 B of spills, 9.5 % of the executed instructions were
// local-memory traffic); the bracket and the foot press in < 3 % of the envs each, so their rows are parked.
template <int N, int NC>
__host__ __device__ constexpr unsigned parked_rows_mask() {
    return (N == 5 && NC == 4) ? 0b1001u : (N == 5 && NC == 3) ? 0b100u : 0u;
}
__host__ __device__ constexpr int popcount_c(unsigned m) { return m ? (int)(m & 1u) + popcount_c(m >> 1) : 0; }

template <int N, int NC>
struct ColdSlots {
    static constexpr int ROWS = N + 3 * NC;
    static constexpr unsigned PARKED = parked_rows_mask<N, NC>();
    static constexpr int LAM = 0;                 // [ROWS]
    static constexpr int MASS = LAM + ROWS;       // [N] mass coefficient
    static constexpr int DAMP = MASS + N;         // [N] damping
    static constexpr int FRIC = DAMP + N;         // [N] friction * dt  (impulse bound)
    static constexpr int MU = FRIC + N;           // [NC]
    static constexpr int TAU = MU + NC;           // [N]
    static constexpr int QLO = TAU + N;           // [N]
    static constexpr int VLO = QLO + N;           // [N]
    static constexpr int CX = VLO + N;            // [3*NC] contact centres (world)
    static constexpr int ACT = CX + 3 * NC;       // [2] this step's clamped action (0 when it was not finite)
    static constexpr int AFIN = ACT + 2;          // [1] 1 = the action was finite
    static constexpr int AOLD = AFIN + 1;         // [2] previous action (loaded in the prologue, used by the epilogue)
    static constexpr int MISC = AOLD + 2;         // [3] episode step counter, episode return (lo, hi words), bit patterns
    static constexpr int QHI = MISC + 3;          // [N] q_hi, v parked after the last iteration: the fp64 epilogue runs
    static constexpr int VHI = QHI + N;           // [N]   once per env and reads everything per half from here
    static constexpr int GROW = VHI + N;          // [parked proxies][3 rows][N + 2]: G_r, reciprocal diagonal, row velocity
    static constexpr int COUNT = GROW + popcount_c(PARKED) * 3 * (N + 2);
    // slot of entry k (0..N-1: G, N: reciprocal diagonal, N+1: row velocity) of row d of parked proxy c
    __host__ __device__ static constexpr int grow(int c, int d, int k) {
        return GROW + (popcount_c(PARKED & ((1u << c) - 1u)) * 3 + d) * (N + 2) + k;
    }
};

template <typename V, int STRIDE>
struct Cold {                    // accessor: slot k of this thread (STRIDE = threads per block, compile time so
    V *base;                     // that every access is base + immediate offset; base = &smem[threadIdx.x])
    using S = typename VT<V>::S;
    __device__ __forceinline__ V &operator()(int k) const { return base[k * STRIDE]; }
    // half h of slot k as a scalar (the epilogue's view: one env at a time)
    __device__ __forceinline__ S &half(int k, int h) const { return reinterpret_cast<S *>(base + k * STRIDE)[h]; }
};

template <typename V, int N>
struct EnvRegs {                 // hot per-thread state
    V q_hi[N], v[N];
    V gz;                        // gravity z (negative)
};

// ------------------------------------------------------------------------------------------------
// structure signature of a model: which entries of its constant tables are exactly 0 / 1, found on the host
// (model_signature in os2r_capi.cu, |x| <= 1e-15 counts as 0). The forward pass is instantiated on it, so that a joint
// whose frame coincides with its parent's (tree_R = 1), differs by a turn about the joint axis x, about y or about z, or whose
// origin / a proxy's centre has zero components, skips the multiplications by those constants (they live in the constant
// bank: the compiler cannot). Joints: 6 bits each, bits 0-2 = kind of tree_R, bits 3-5 = non-zero components of tree_p;
// bits 30-31: 0 = whether the root body is folded (ModelDev::root_spin) is read at run time, 1 = it is not, 2 = it is.
// Proxies: 6 bits each, bits 0-2 = non-zero components of contact_pos, bits 3-5 = the body that carries the proxy
// (7 = read ModelDev::contact_body at run time). Known at compile time, the root fold and the proxy -> body map cost no
// warp-uniform branches and leave no dead code in the loop body (~3.5 KB of the 38 KB the 4-joint model's loop had,
// against a 32 KB instruction cache). The all-general signature runs any model.
// ------------------------------------------------------------------------------------------------
enum { OS2R_TREE_GENERAL = 0, OS2R_TREE_IDENTITY = 1, OS2R_TREE_XTURN = 2, OS2R_TREE_ZTURN = 3, OS2R_TREE_YTURN = 4 };
__host__ __device__ constexpr uint32_t generic_joint_signature(int n) {
    return n ? ((generic_joint_signature(n - 1) << 6) | (7u << 3) | OS2R_TREE_GENERAL) : 0u;
}
__host__ __device__ constexpr uint32_t generic_contact_signature(int nc) {
    return nc ? ((generic_contact_signature(nc - 1) << 6) | (7u << 3) | 7u) : 0u;
}
// proxy k rides on body i / the root body is folded: compile-time constants when the signature knows
template <uint32_t SC, typename MT>
__device__ __forceinline__ bool proxy_on_body(const MT &M, int k, int i) {
    const uint32_t b = (SC >> (6 * k + 3)) & 7u;
    return b == 7u ? M.contact_body[k] == i : (int)b == i;
}
template <uint32_t SC, typename MT>
__device__ __forceinline__ bool joint_moves_proxy(const MT &M, int k, int i) {
    const uint32_t b = (SC >> (6 * k + 3)) & 7u;
    return b == 7u ? i <= M.contact_body[k] : i <= (int)b;
}
template <uint32_t SJ, typename MT>
__device__ __forceinline__ bool root_folded(const MT &M) {
    return (SJ >> 30) == 0u ? M.root_spin != 0 : (SJ >> 30) == 2u;
}

template <typename V, int N, int NC, bool DAMPED, uint32_t SJ, uint32_t SC, typename ColdT>
__device__ __forceinline__ void physics_iteration(const ModelDev<typename VT<V>::S> &M, EnvRegs<V, N> &E, const ColdT &C) {
    using SL = ColdSlots<N, NC>;
    using T = typename VT<V>::S;
    using Mask = typename VT<V>::M;
    constexpr bool FP32 = sizeof(T) == 4;
    const T dt = M.dt;

    // ---- single forward pass: kinematics, velocities, and direct accumulation of the joint-space mass
    //      matrix  M = sum_b [ m Jv^T Jv + Jw^T Iw Jw ]  and bias  h = sum_b [ Jv^T m a_c + Jw^T (Iw al + w x Iw w) ]
    //      (classical world-frame Newton-Euler with qdd = 0; gravity as an upward base acceleration).
    //      Every term is a difference of nearby points or a positive contribution: no reference point,
    //      no m|r|^2 cancellation, and nothing per-body has to be kept for a backward pass.
    V ax[N][3];      // joint axes, world
    V P[N][3];       // joint origins, world
    V Mm[N][N];      // lower triangle (Mm[j][k], k <= j)
    V hb[N];         // bias forces
#pragma unroll
    for (int j = 0; j < N; ++j) {
        hb[j] = V(0);
#pragma unroll
        for (int k = 0; k <= j; ++k) Mm[j][k] = V(0);
    }
    V sn[N], cs[N];      // sin / cos of every joint angle (hi + lo)
    if constexpr (FP32) {
        bool big = false;
#pragma unroll
        for (int i = 0; i < N; ++i) big = big || beyond_fast_range(E.q_hi[i]);
        if (__builtin_expect(big, 0)) {       // a joint that has turned > 15 000 revolutions (or is non-finite): library range reduction
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const SinCos<V> sc = slow_sincos(E.q_hi[i], C(SL::QLO + i));
                sn[i] = sc.s; cs[i] = sc.c;
            }
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                V s, c;
                sincos_fast<V>(E.q_hi[i], &s, &c);
                const V lo = C(SL::QLO + i);
                sn[i] = s + lo * c;
                cs[i] = c - lo * s;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const SinCos<V> sc = joint_sincos<V>(E.q_hi[i], V(0));
            sn[i] = sc.s; cs[i] = sc.c;
        }
    }
    {
        V R[9] = {V(1), V(0), V(0), V(0), V(1), V(0), V(0), V(0), V(1)};
        V p[3] = {V(0), V(0), V(0)};
        V w[3] = {V(0), V(0), V(0)}, al[3] = {V(0), V(0), V(0)}, ap[3] = {V(0), V(0), -E.gz};   // ang. vel, ang. acc, acc of joint origin
#pragma unroll
        for (int i = 0; i < N; ++i) {
            V A[9];
            if (i == 0) {
#pragma unroll
                for (int k = 0; k < 3; ++k) p[k] = V(M.tree_p[0][k]);
#pragma unroll
                for (int k = 0; k < 9; ++k) A[k] = V(M.tree_R[0][k]);
            } else {
                const uint32_t tree_kind = (SJ >> (6 * i)) & 7u, p_mask = (SJ >> (6 * i + 3)) & 7u;   // constants after unrolling
                const T *tp = M.tree_p[i];
                if (p_mask) {
                    V dp[3], wd[3];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        V acc = V(0);
                        bool first = true;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            if ((p_mask >> c) & 1u) { acc = first ? R[3 * r + c] * tp[c] : fma_t(R[3 * r + c], V(tp[c]), acc); first = false; }
                        dp[r] = acc;
                    }
                    // acceleration of the next joint origin, carried by the parent body: ap += al x dp + w x (w x dp)
                    OS2R_CROSS(wd, w, dp);
                    OS2R_CROSS_ACC(ap, al, dp);
                    OS2R_CROSS_ACC(ap, w, wd);
#pragma unroll
                    for (int r = 0; r < 3; ++r) p[r] += dp[r];
                }
                const T *tR = M.tree_R[i];
                if (tree_kind == OS2R_TREE_GENERAL) {
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            A[3 * r + c] = R[3 * r] * tR[c] + R[3 * r + 1] * tR[3 + c] + R[3 * r + 2] * tR[6 + c];
                } else if (tree_kind == OS2R_TREE_ZTURN) {      // tree_R = turn about z: third column unchanged
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        A[3 * r] = R[3 * r] * tR[0] + R[3 * r + 1] * tR[3];
                        A[3 * r + 1] = R[3 * r] * tR[1] + R[3 * r + 1] * tR[4];
                        A[3 * r + 2] = R[3 * r + 2];
                    }
                } else if (tree_kind == OS2R_TREE_YTURN) {      // tree_R = turn about y: second column unchanged
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        A[3 * r] = R[3 * r] * tR[0] + R[3 * r + 2] * tR[6];
                        A[3 * r + 1] = R[3 * r + 1];
                        A[3 * r + 2] = R[3 * r] * tR[2] + R[3 * r + 2] * tR[8];
                    }
                } else {                                         // identity, or a turn about x (folded into s, c below)
#pragma unroll
                    for (int k = 0; k < 9; ++k) A[k] = R[k];
                }
            }
            V s = sn[i], c = cs[i];
            if (i > 0 && ((SJ >> (6 * i)) & 7u) == OS2R_TREE_XTURN) {
                // tree_R = turn about the joint axis by phi: R Rx(phi) Rx(q) = R Rx(phi + q) — the angle addition on
                // (sin, cos) with the table's own cos phi = tree_R[4], sin phi = tree_R[7] replaces the 3x3 product
                const T cphi = M.tree_R[i][4], sphi = M.tree_R[i][7];
                const V c2 = c * cphi - s * sphi, s2 = s * cphi + c * sphi;
                c = c2; s = s2;
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const V a1 = A[3 * r + 1], a2 = A[3 * r + 2];
                R[3 * r] = A[3 * r];
                R[3 * r + 1] = c * a1 + s * a2;
                R[3 * r + 2] = c * a2 - s * a1;
                ax[i][r] = A[3 * r];
                P[i][r] = p[r];
            }
            const V qd = E.v[i];
            if (i > 0) {   // al += (w_parent x a_i) qd
                V wa[3];
                OS2R_CROSS(wa, w, ax[i]);
#pragma unroll
                for (int r = 0; r < 3; ++r) al[r] += wa[r] * qd;
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) w[r] += ax[i][r] * qd;
            if (i == 0 && root_folded<SJ>(M)) {   // ~130 instructions of body 0 become one FMA
                Mm[0][0] = fma_t(V(M.root_mass_term), C(SL::MASS), V(M.root_inertia_term));
            } else {
            // COM offset and rotational inertia in world axes
            const T *cm = M.com[i];
            V d[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) d[r] = R[3 * r] * cm[0] + R[3 * r + 1] * cm[1] + R[3 * r + 2] * cm[2];
            V Iw[6];
            {
                const T *Ib = M.inertia[i];
                V t[9];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    t[3 * r + 0] = R[3 * r] * Ib[0] + R[3 * r + 1] * Ib[3] + R[3 * r + 2] * Ib[4];
                    t[3 * r + 1] = R[3 * r] * Ib[3] + R[3 * r + 1] * Ib[1] + R[3 * r + 2] * Ib[5];
                    t[3 * r + 2] = R[3 * r] * Ib[4] + R[3 * r + 1] * Ib[5] + R[3 * r + 2] * Ib[2];
                }
                Iw[0] = t[0] * R[0] + t[1] * R[1] + t[2] * R[2];
                Iw[1] = t[3] * R[3] + t[4] * R[4] + t[5] * R[5];
                Iw[2] = t[6] * R[6] + t[7] * R[7] + t[8] * R[8];
                Iw[3] = t[0] * R[3] + t[1] * R[4] + t[2] * R[5];
                Iw[4] = t[0] * R[6] + t[1] * R[7] + t[2] * R[8];
                Iw[5] = t[3] * R[6] + t[4] * R[7] + t[5] * R[8];
            }
            const V m = C(SL::MASS + i) * M.mass[i];
            // body wrench about its COM (qdd = 0): f = m (ap + al x d + w x (w x d)), n = Iw al + w x (Iw w)
            V f[3], nn[3], wd[3], Iwv[3];
            OS2R_CROSS(wd, w, d);
#pragma unroll
            for (int r = 0; r < 3; ++r) f[r] = ap[r];
            OS2R_CROSS_ACC(f, al, d);
            OS2R_CROSS_ACC(f, w, wd);
#pragma unroll
            for (int r = 0; r < 3; ++r) f[r] *= m;
            OS2R_SYMV(Iwv, Iw, w);
            OS2R_SYMV(nn, Iw, al);
            OS2R_CROSS_ACC(nn, w, Iwv);
            // joint-space accumulation over this body's ancestors j <= i
            V Jv[N][3], u[N][3];
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const V rr[3] = {p[0] + d[0] - P[j][0], p[1] + d[1] - P[j][1], p[2] + d[2] - P[j][2]};
                OS2R_CROSS(Jv[j], ax[j], rr);
                OS2R_SYMV(u[j], Iw, ax[j]);
                OS2R_DOT_ACC(hb[j], Jv[j], f);
                OS2R_DOT_ACC(hb[j], ax[j], nn);
#pragma unroll
                for (int k = 0; k <= j; ++k) {
                    Mm[j][k] = fma_t(m, OS2R_DOT(Jv[j], Jv[k]), Mm[j][k]);
                    OS2R_DOT_ACC(Mm[j][k], ax[j], u[k]);
                }
            }
            }
            // contact spheres carried by this body
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                if (proxy_on_body<SC>(M, k, i)) {
                    const T *cp = M.contact_pos[k];
                    const uint32_t c_mask = (SC >> (6 * k)) & 7u;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        V acc = p[r];
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc)
                            if ((c_mask >> cc) & 1u) acc = fma_t(R[3 * r + cc], V(cp[cc]), acc);
                        C(SL::CX + 3 * k + r) = acc;
                    }
                }
            }
        }
    }
    if (!(OS2R_SKIP_PHASE_BARRIERS & 2)) __syncthreads();   // second phase-alignment point per iteration (see the note at the physics loop)
    // ---- Cholesky of M (and of M + dt*D when any joint is damped); qdd; v* = v + dt*qdd ------------------
    V L[N][N];       // Cholesky factor of the plain M (lower); Ld = reciprocal diagonal
    V Ld[N];
    V vs[N], dvq[N];
    auto cholesky = [&](const V *dadd, V(&Lo)[N][N], V(&Lrd)[N]) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            V d = Mm[j][j] + dadd[j];
#pragma unroll
            for (int k = 0; k < j; ++k) d -= Lo[j][k] * Lo[j][k];
            const V rd = rsqrt_t(d);
            Lo[j][j] = d * rd;
            Lrd[j] = rd;
#pragma unroll
            for (int i = j + 1; i < N; ++i) {
                V s = Mm[i][j];
#pragma unroll
                for (int k = 0; k < j; ++k) s -= Lo[i][k] * Lo[j][k];
                Lo[i][j] = s * rd;
            }
        }
    };
    auto solve = [&](const V(&Lo)[N][N], const V(&Lrd)[N], const V *b, V *x) {
        V y[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            V s = b[i];
#pragma unroll
            for (int k = 0; k < i; ++k) s -= Lo[i][k] * y[k];
            y[i] = s * Lrd[i];
        }
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            V s = y[i];
#pragma unroll
            for (int k = i + 1; k < N; ++k) s -= Lo[k][i] * x[k];
            x[i] = s * Lrd[i];
        }
    };
    {
        V rhs[N], zero[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            zero[i] = V(0);
            rhs[i] = C(SL::TAU + i) - C(SL::DAMP + i) * E.v[i] - hb[i];
        }
        cholesky(zero, L, Ld);
        V qdd[N];
        if (DAMPED) {   // compile-time: the second factorisation is ~1.6 KB of loop body the undamped models never run
            V dd[N], L2[N][N], L2d[N];
#pragma unroll
            for (int i = 0; i < N; ++i) dd[i] = C(SL::DAMP + i) * dt;
            cholesky(dd, L2, L2d);
            solve(L2, L2d, rhs, qdd);
        } else {
            solve(L, Ld, rhs, qdd);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            dvq[i] = qdd[i] * dt;
            vs[i] = E.v[i] + dvq[i];
        }
    }
    // whitened velocity z0 = L^T v*. The solver tracks only the impulse-induced change z (total = z0 + z):
    // v_new = v* + L^-T z, so the (usually tiny) constraint correction never round-trips v through L.
    V z0[N], z[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        V s = V(0);
#pragma unroll
        for (int k = i; k < N; ++k) s += L[k][i] * vs[k];
        z0[i] = s;
        z[i] = V(0);
    }
    // ---- constraint rows in whitened coordinates: G_r = L^-1 J_r^T ----------------------------------------
    V Gj[N][N];      // joint friction row r: column r of L^-1 (entries k >= r)
    V Aj[N], bj[N];  // reciprocal regularised diagonal; row velocity before impulses
#pragma unroll
    for (int r = 0; r < N; ++r) {
        V a = V(0), b = V(0);
#pragma unroll
        for (int k = r; k < N; ++k) {
            V s = (k == r) ? V(1) : V(0);
#pragma unroll
            for (int m = r; m < k; ++m) s -= L[k][m] * Gj[r][m];
            Gj[r][k] = s * Ld[k];
            a += Gj[r][k] * Gj[r][k];
            b += Gj[r][k] * z0[k];
        }
        Aj[r] = rcp_t(a * M.cfm1_joint);
        bj[r] = b;
        // warm start; a row without friction (bound 0) starts from 0 and, its bound being 0, stays there
        const V l = sel_t(gt_t(C(SL::FRIC + r), V(0)), C(SL::LAM + r), V(0));
#pragma unroll
        for (int k = r; k < N; ++k) z[k] += Gj[r][k] * l;
        C(SL::LAM + r) = l;
    }
    if (!(OS2R_SKIP_PHASE_BARRIERS & 4)) __syncthreads();   // third phase-alignment point: the contact phase starts together (88.2 -> 87.4 us per step)
    V Gc[NC][3][N];
    V Ac[NC][3];
    V bc[NC][3];     // row velocity before impulses, minus the target (penetration correction)
    bool act[NC];    // some env of this thread presses proxy c into the ground
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const V cz = C(SL::CX + 3 * c + 2);
        const V depth = V(M.contact_radius[c]) - cz;
        const Mask on_c = gt_t(depth, V(0));
        act[c] = any_t(on_c);
        // every proxy but the hip sphere (index NC - 3 in the shipped models) presses in < 6 % of the envs: their rows are
        // marked unlikely so that ptxas lays them out behind the loop body (the hot path stays contiguous for the
        // instruction cache) — placement only, the arithmetic is the same
        if ((c == NC - 3) ? act[c] : __builtin_expect(act[c], 0)) {
            const V bounce = fmin_t(depth * M.erp_over_dt, V(M.max_erv));
            const V x[3] = {C(SL::CX + 3 * c), C(SL::CX + 3 * c + 1), cz - M.contact_radius[c]};   // lowest point
            V J[3][N];   // rows: normal (z), tangent x, tangent y
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const V rr[3] = {x[0] - P[i][0], x[1] - P[i][1], x[2] - P[i][2]};
                V jc[3];
                OS2R_CROSS(jc, ax[i], rr);
                const bool on = joint_moves_proxy<SC>(M, c, i);
                J[0][i] = on ? jc[2] : V(0);
                J[1][i] = on ? jc[0] : V(0);
                J[2][i] = on ? jc[1] : V(0);
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                V a = V(0), b = (d == 0) ? -bounce : V(0);
                V g[N];
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    V s = J[d][k];
#pragma unroll
                    for (int m = 0; m < k; ++m) s -= L[k][m] * g[m];
                    g[k] = s * Ld[k];
                    a += g[k] * g[k];
                    b += g[k] * z0[k];
                }
                // An env of the pair that is NOT in contact gets a dead row: relaxation factor, row velocity and
                // impulse 0, so every update below computes lam' = 0, dlam = 0 for it (a single-env build never
                // gets here for such an env: the selects are no-ops).
                const V ra = sel_t(on_c, rcp_t(a * M.cfm1_contact), V(0)), rb = sel_t(on_c, b, V(0));
                const V l = sel_t(on_c, C(SL::LAM + N + 3 * c + d), V(0));   // warm start
#pragma unroll
                for (int k = 0; k < N; ++k) z[k] += g[k] * l;
                C(SL::LAM + N + 3 * c + d) = l;
                if ((SL::PARKED >> c) & 1u) {          // compile-time after unrolling: rows of a rarely pressed proxy
#pragma unroll
                    for (int k = 0; k < N; ++k) C(SL::grow(c, d, k)) = g[k];
                    C(SL::grow(c, d, N)) = ra;
                    C(SL::grow(c, d, N + 1)) = rb;
                } else {
#pragma unroll
                    for (int k = 0; k < N; ++k) Gc[c][d][k] = g[k];
                    Ac[c][d] = ra;
                    bc[c][d] = rb;
                }
            }
        } else {
#pragma unroll
            for (int d = 0; d < 3; ++d) C(SL::LAM + N + 3 * c + d) = V(0);
        }
    }
    // ---- projected Gauss-Seidel sweeps in whitened coordinates -------------------------------------------
    // Row update with relative CFM c on the diagonal A(1+c):
    //   lam' = clamp(lam - (G.z - target + c*A*lam) / (A(1+c))) = clamp(lam*(1-k) - (G.z - target)*inv),
    //   1-k = 1/(1+c), inv = 1/(A(1+c)) precomputed per row.
    // Early exit (per env): the sweeps of an env end after the first sweep whose whitened velocity change
    // |dz| = sqrt(dv^T M dv) is <= pgs_tol (tol 0: only when the sweep left z bit-for-bit unchanged, the typical
    // case being saturated joint friction without contact). The decision uses the env's own data only, so a
    // result never depends on which other envs share the thread or the warp: in the pair build an env that has
    // finished is FROZEN (its rows keep their impulses, dlam = 0) while its partner sweeps on; the warp leaves the
    // loop when its last env does.
    const T kj1 = M.kj1, kc1 = M.kc1;   // 1 - k
    const V tol2 = V(M.pgs_tol2);
    Mask live = all_true<V>();
#pragma unroll 1
    for (int it = 0; it < M.pgs_iters; ++it) {
        V zs[N];
#pragma unroll
        for (int k = 0; k < N; ++k) zs[k] = z[k];
        if (M.pgs_joint_sweeps <= 0 || it < M.pgs_joint_sweeps)   // os2r_model.pgs_joint_sweeps (warp-uniform)
#pragma unroll
        for (int r = 0; r < N; ++r) {
            // branch-free: a row without friction has bound 0, so its impulse stays 0 and the update adds 0
            const V lam = C(SL::LAM + r), lim = C(SL::FRIC + r);
            V w = bj[r];
#pragma unroll
            for (int k = r; k < N; ++k) w += Gj[r][k] * z[k];
            V nl = lam * kj1 - w * Aj[r];
            nl = fmax_t(-lim, fmin_t(lim, nl));
            if constexpr (VT<V>::LANES > 1) nl = sel_t(live, nl, lam);
            const V dl = nl - lam;
#pragma unroll
            for (int k = r; k < N; ++k) z[k] += Gj[r][k] * dl;
            C(SL::LAM + r) = nl;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if ((c == NC - 3) ? act[c] : __builtin_expect(act[c], 0)) {
                V ln = C(SL::LAM + N + 3 * c);
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const int r = N + 3 * c + d;
                    const V lam = (d == 0) ? ln : C(SL::LAM + r);
                    V g[N], ra, w;
                    if ((SL::PARKED >> c) & 1u) {      // compile-time after unrolling
#pragma unroll
                        for (int k = 0; k < N; ++k) g[k] = C(SL::grow(c, d, k));
                        ra = C(SL::grow(c, d, N));
                        w = C(SL::grow(c, d, N + 1));
                    } else {
#pragma unroll
                        for (int k = 0; k < N; ++k) g[k] = Gc[c][d][k];
                        ra = Ac[c][d];
                        w = bc[c][d];
                    }
#pragma unroll
                    for (int k = 0; k < N; ++k) w += g[k] * z[k];
                    V nl = lam * kc1 - w * ra;
                    if (d == 0) nl = fmax_t(nl, V(0));
                    else {
                        const V lim = C(SL::MU + c) * ln;
                        nl = fmax_t(-lim, fmin_t(lim, nl));
                    }
                    if constexpr (VT<V>::LANES > 1) nl = sel_t(live, nl, lam);
                    if (d == 0) ln = nl;
                    const V dl = nl - lam;
#pragma unroll
                    for (int k = 0; k < N; ++k) z[k] += g[k] * dl;
                    C(SL::LAM + r) = nl;
                }
            }
        }
        V e2 = V(0);
#pragma unroll
        for (int k = 0; k < N; ++k) { const V dz = z[k] - zs[k]; e2 += dz * dz; }
        live = and_not(live, le_t(e2, tol2));
        if (!any_t(live)) break;
    }
    // ---- v = v* + L^-T z ; q += dt v  (TwoSum-compensated (hi, lo) pairs in fp32) --------------------------
    {
        V dv[N];
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            V s = z[i];
#pragma unroll
            for (int k = i + 1; k < N; ++k) s -= L[k][i] * dv[k];
            dv[i] = s * Ld[i];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if constexpr (FP32) {
                {   // velocity: measured 10x less drift over 1000 steps than plain fp32 accumulation
                    const V b = (dvq[i] + dv[i]) + C(SL::VLO + i);
                    const V a = E.v[i];
                    const V s = a + b;
                    const V bb = s - a;
                    C(SL::VLO + i) = (a - (s - bb)) + (b - bb);
                    E.v[i] = s;
                }
                {   // position
                    const V b = E.v[i] * dt + C(SL::QLO + i);
                    const V a = E.q_hi[i];
                    const V s = a + b;
                    const V bb = s - a;
                    C(SL::QLO + i) = (a - (s - bb)) + (b - bb);
                    E.q_hi[i] = s;
                }
            } else {
                E.v[i] = vs[i] + dv[i];
                E.q_hi[i] += E.v[i] * dt;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// task epilogue (fp64): observation, reward, termination — once per env step
// ------------------------------------------------------------------------------------------------
__device__ inline double np_mod_pos(double x, double y) {
    double r = fmod(x, y);
    if (r != 0.0) { if (r < 0.0) r += y; } else r = 0.0;
    return r;
}

// raw[] (masked, wrapped) and normalised obs[]; returns done-by-task.
// FAST (the fp32 product build; the observation leaves the device as float32 either way): position columns are
// normalised with the host's 2 / (high - low) in one fp64 FMA instead of an fp64 division, velocity columns through the
// fp32 tanhf (<= 2 ulp of a float) instead of the fp64 tanh — together 8 % of the warp-time of a contact-free step
// (profiles/r2_step_kernel_fresh_by_region.txt). Termination is decided on the RAW fp64 values in both builds.
template <int N, bool FAST = false>
__device__ inline bool observe(const TaskDev &K, const double *q, const double *v, const double *a_old,
                               double *obs) {
    const os2r_task_cfg &C = K.cfg;
    const double PI = 3.141592653589793;
    bool done = false;
#pragma unroll 1
    for (int k = 0; k < C.obs_dim; ++k) {
        const int kind = C.obs_kind[k], idx = C.obs_index[k];
        double x = 0;
        // idx is runtime: select without dynamic register indexing
        if (kind == OS2R_OBS_TORQUE) x = idx == 0 ? a_old[0] : a_old[1];
        else {
#pragma unroll
            for (int i = 0; i < N; ++i)
                if (i == idx) x = (kind == OS2R_OBS_VEL) ? v[i] : q[i];
        }
        if (kind == OS2R_OBS_POS_PERIODIC) x = np_mod_pos(x + PI, 2 * PI) - PI;
        done |= !(x >= C.done_low[k]) || !(x <= C.done_high[k]);
        double o = x;
        if (C.normalized) {
            if (kind == OS2R_OBS_VEL) o = FAST ? (double)tanhf((float)(0.05 * x)) : tanh(0.05 * x);
            else o = FAST ? fma(x - C.obs_low[k], K.obs_scale[k], -1.0) : 2 * (x - C.obs_low[k]) / (C.obs_high[k] - C.obs_low[k]) - 1;
        }
        obs[k] = o;
    }
    return done;
}

__device__ inline double quad_tol(double x, double margin, double vam) {   // tolerance(x, (0,0), margin, quadratic)
    if (x == 0.0) return 1.0;
    const double sx = (fabs(x) / margin) * sqrt(1 - vam);
    return fabs(sx) < 1 ? 1 - sx * sx : 0.0;
}
__device__ inline double lin_tol(double x, double margin, double vam) {
    if (x == 0.0) return 1.0;
    const double sx = (fabs(x) / margin) * (1 - vam);
    return fabs(sx) < 1 ? 1 - sx : 0.0;
}

// rewards/__init__.py:66-207 ; a0 current action, a1 previous
__device__ inline double reward_fn(const os2r_task_cfg &C, const double *obs, const double *a0, const double *a1) {
    const double H = C.normalized ? 0.11 / 1.57 : 0.11;
    double bp = 0;
    if (C.reward_pitch_col >= 0) bp = obs[C.reward_pitch_col];
    const double band = (H <= bp && bp <= 4 * H) ? 1.0 : 0.0;
    switch (C.reward_id) {
    case OS2R_REWARD_BALANCING_V1: return band;
    case OS2R_REWARD_BALANCING_V2: return band * quad_tol(a0[0], 1, 0.4) * quad_tol(a0[1], 1, 0.4);
    case OS2R_REWARD_BALANCING_V3: {
        double up = 1.0;
        if (band == 0.0) {
            const double d = (bp < H ? H - bp : bp - 4 * H) / 0.01;
            const double sc = sqrt(1 / 0.1 - 1);
            up = 1 / ((d * sc) * (d * sc) + 1);
        }
        return up * quad_tol(a0[0] - a1[0], 1, 0.1) * quad_tol(a0[1] - a1[1], 1, 0.1); }
    case OS2R_REWARD_HOPPING_V1: {
        const double hv = obs[C.reward_yawvel_col];
        double move = 1.0;
        if (!(0.25 <= hv && hv <= 0.3)) {
            const double d = (hv < 0.25 ? 0.25 - hv : hv - 0.3) / 0.15;
            const double t = tanh(d * atanh(sqrt(1 - 0.1)));
            move = 1 - t * t;
        }
        return band * quad_tol(a0[0] - a1[0], 0.1, 0.0) * quad_tol(a0[1] - a1[1], 0.1, 0.0) * move; }
    case OS2R_REWARD_STRAIGHT_V1: {
        double sc = (quad_tol(a0[0] / 20, 1, 0.0) + quad_tol(a0[1] / 20, 1, 0.0)) / 2;
        sc = (4 + sc) / 5;
        return lin_tol(obs[C.reward_hip_col], 1, 0.1) * lin_tol(obs[C.reward_knee_col], 1, 0.1) * sc; }
    default: return 0.0;
    }
}

// utils/reset.py:4-40
__device__ inline void leg_joint_angles(const os2r_task_cfg &C, double bp, double out[2]) {
    const double lh = (C.ik_boom * sin(bp) + C.ik_pivot_height) / cos(bp);
    const double ul = C.ik_upper_leg, ll = C.ik_lower_leg;
    const double lleg = lh - C.ik_hip_offset - C.ik_clip;
    if (lleg > ul + ll) { out[0] = 0; out[1] = 0; return; }
    double ca = (ul * ul + lleg * lleg - ll * ll) / (2 * ul * lleg);
    ca = fmin(1.0, fmax(-1.0, ca));
    const double hip = acos(ca);
    double sa = ul * sin(hip) / ll;
    sa = fmin(1.0, fmax(-1.0, sa));
    out[0] = hip;
    out[1] = -(asin(sa) + hip);
}

// per-env parameter draws for episode `ep` (randomizers/monopod.py:182-215); writes SoA params
template <typename T, typename D>
__device__ inline void draw_params(const TaskDev &K, StateDev<T> &S, int64_t e, uint32_t ep, const D &draw) {
    const os2r_task_cfg &C = K.cfg;
    const int64_t N = S.n_envs;
    for (int i = 0; i < K.n_dof; ++i) {
        double ms = 1.0, dm = K.nominal_damping[i], fr = K.nominal_friction[i];
        if (C.randomize_params) {
            ms = C.mass_lo + (C.mass_hi - C.mass_lo) * draw(DRAW_PARAMS + i);
            fr = C.fric_lo + (C.fric_hi - C.fric_lo) * draw(DRAW_PARAMS + OS2R_MAX_DOF + i);
            dm = K.nominal_damping[i] * (C.damp_lo + (C.damp_hi - C.damp_lo) * draw(DRAW_PARAMS + 2 * OS2R_MAX_DOF + i));
        }
        S.mass_scale[i * N + e] = (T)ms;
        S.damping[i * N + e] = (T)dm;
        S.friction[i * N + e] = (T)fr;
    }
    for (int c = 0; c < K.n_contacts; ++c) {
        double mu = K.nominal_mu[c];
        if (C.randomize_params)
            mu = C.mu_link * (C.mu_lo + (C.mu_hi - C.mu_lo) * draw(DRAW_PARAMS + 3 * OS2R_MAX_DOF + c));
        S.mu[c * N + e] = (T)mu;
    }
    // MonopodEnvRandomizer(num_physics_rollouts=K): randomize_physics again at every K-th reset of the env
    // (randomizers/monopod.py:36,56-61,371)
    if (C.randomize_gravity && C.gravity_redraw_resets > 0 && ep % (uint32_t)C.gravity_redraw_resets == 0) {
        double z[2];
        draw_normal2(draw, DRAW_GRAVITY, z);
        S.gravity_z[e] = (T)(C.grav_mean + C.grav_std * z[0]);
    }
}

// Reset pose draw (randomizers/monopod.py:67-135, monopod_no_rand.py:26-98). q[] in chain order.
template <int N, typename D>
__device__ inline int reset_pose(const TaskDev &K, const D &draw, double *q) {
    const os2r_task_cfg &C = K.cfg;
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = 0;
    int idx = (int)(draw(DRAW_RESET) * C.n_resets);
    if (idx >= C.n_resets) idx = C.n_resets - 1;
    double pitch = C.reset_pitch[idx];
    double leg[2];
    double yaw = 0;
    if (C.reset_randomized) {
        pitch *= 0.8 + 0.4 * draw(DRAW_PITCH);
        double zz[2];
        draw_normal2(draw, DRAW_NOISE, zz);
        const double r0 = fabs(0.2 * zz[0]), r1 = fabs(0.2 * zz[1]);
        const double rmax = r0 > r1 ? r0 : r1, rmin = r0 > r1 ? r1 : r0;
        if (!C.reset_laying[idx]) leg_joint_angles(C, pitch, leg);
        else { leg[0] = 1.57 - (draw(DRAW_LAYSIDE) < 0.5 ? 3.14 : 0.0); leg[1] = 0; }
        leg[0] = leg[0] + (leg[0] > 0 ? 1.0 : 0.0) * rmax;   // (a>0 - a<0) == (a>0): reference precedence quirk
        leg[1] = leg[1] - (leg[1] > 0 ? 1.0 : 0.0) * rmin;
        const double dir = 1.0 - (draw(DRAW_DIR) < 0.5 ? 2.0 : 0.0);
        leg[0] *= dir; leg[1] *= dir;
        yaw = -0.2 + 0.4 * draw(DRAW_YAW);
    } else if (C.simple_sample_reset) {
        leg[0] = C.simple_lo[0] + (C.simple_hi[0] - C.simple_lo[0]) * draw(DRAW_SIMPLE_HIP);
        leg[1] = C.simple_lo[1] + (C.simple_hi[1] - C.simple_lo[1]) * draw(DRAW_SIMPLE_KNEE);
    } else {
        if (!C.reset_laying[idx]) leg_joint_angles(C, pitch, leg);
        else { leg[0] = 1.57; leg[1] = 0; }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (i == K.role_dof[OS2R_ROLE_YAW]) q[i] = yaw;
        if (i == K.role_dof[OS2R_ROLE_PITCH]) q[i] = pitch;
        if (i == K.role_dof[OS2R_ROLE_HIP]) q[i] = leg[0];
        if (i == K.role_dof[OS2R_ROLE_KNEE]) q[i] = leg[1];
    }
    return idx;
}

}  // namespace os2r
