/*
 * os2r_oracle.c — CPU restatement (fp64, plain C) of the monopod step path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product (gym_os2r_b200/) imports, links or calls this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY STATUS
 *   - task logic (observation, normalisation, reward, termination, reset-pose IK): follows the
 *     reference's own numpy code line by line and is PINNED by golden vectors generated from that
 *     code (tests/golden/task_kat.json, made by tools/gen_golden.py importing /root/reference).
 *   - physics: PARITY UNPINNED. The arithmetic lives in third-party, un-vendored, unpinned
 *     dependencies (setup.py:22-26: gym-ignition -> ScenarIO -> Ignition Gazebo -> ign-physics
 *     dartsim -> DART 6.x) that are absent from /root/reference and from this image, and the
 *     reference's tests hold no trajectory / golden vector for it (tests/tests_general.py).
 *     This file restates DART's published World::step time-stepping scheme:
 *       1. articulated-body forward dynamics with joint damping treated implicitly
 *          (d_i*dt added to the projected articulated inertia; damping force -d_i*qd_i),
 *       2. qd += dt*qdd,
 *       3. constraints at the *current* positions: per-DoF Coulomb joint friction rows
 *          (impulse bound +-friction*dt, target velocity 0) and, per penetrating contact, one
 *          normal row (lambda >= 0, penetration correction min(depth*ERP/dt, MAX_ERV), ERP=0.01,
 *          MAX_ERV=1e-3, CFM=1e-5 relative) and two friction rows on the world x / y axes with the
 *          pyramid bound +-mu*lambda_n, solved as one boxed LCP,
 *       4. qd += Minv J^T lambda (plain, non-implicit inertia), 5. q += dt*qd  (semi-implicit).
 *     The LCP is solved by projected Gauss-Seidel with a FIXED sweep count and warm start from
 *     the previous iteration's impulses (DART: Dantzig pivoting, PGS fallback) — the documented
 *     solver difference of BASELINE.json's north_star. `oracle_substep_converged` runs the same
 *     sweeps to convergence to measure that truncation error. Collision geometry: analytic spheres
 *     (reference: STL trimesh vs plane).
 *   Call sites restated: runtimes/gazebo_runtime.py:65-97 (10x zero-order-hold torque + run()),
 *   tasks/monopod.py:202-236 (torque map), :238-272 (observation), :274-298 (done),
 *   rewards/__init__.py:66-207 + rewards/rewards_utils.py:10-122 (rewards),
 *   randomizers/monopod.py:56-61,67-135,182-215 and monopod_no_rand.py:26-98 (resets / draws),
 *   utils/reset.py:4-40 (IK), common/vec_env/subproc_vec_env.py:14-21 (auto-reset).
 *
 * The dynamics formulation here (body-coordinate Featherstone ABA, Minv from unit-torque ABA
 * solves, as DART's impulse pass does) is deliberately DIFFERENT from the CUDA kernel's
 * (world-aligned composite-rigid-body + Cholesky), so agreement is a real cross-check.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/os2r.h"

#define NMAX OS2R_MAX_DOF
#define CMAX OS2R_MAX_CONTACTS
#define RMAX OS2R_MAX_ROWS

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011) — the counter-based stream shared with the CUDA path.     */
/* key = (seed lo, seed hi); counter = (env id lo, env id hi, episode, block).                  */
/* ------------------------------------------------------------------------------------------ */
void oracle_philox(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t block, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)env_id, c1 = (uint32_t)(env_id >> 32), c2 = episode, c3 = block;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* k-th uniform double in [0,1) of the (env, episode) stream: 53 bits from two words. */
static double rng_uniform(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t k) {
    uint32_t x[4];
    oracle_philox(seed, env_id, episode, k >> 1, x);
    uint32_t a = x[(k & 1) * 2], b = x[(k & 1) * 2 + 1];
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
double oracle_uniform(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t k) {
    return rng_uniform(seed, env_id, episode, k);
}
/* Box-Muller pair from uniforms k, k+1 */
static void rng_normal2(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t k, double z[2]) {
    double u1 = rng_uniform(seed, env_id, episode, k), u2 = rng_uniform(seed, env_id, episode, k + 1);
    double r = sqrt(-2.0 * log(1.0 - u1));
    z[0] = r * cos(6.283185307179586476925 * u2);
    z[1] = r * sin(6.283185307179586476925 * u2);
}
/* draw indices (shared with the CUDA kernels) */
enum { DRAW_RESET = 0, DRAW_PITCH = 1, DRAW_NOISE = 2 /*,3*/, DRAW_LAYSIDE = 4, DRAW_DIR = 5, DRAW_YAW = 6,
       DRAW_SIMPLE_HIP = 7, DRAW_SIMPLE_KNEE = 8, DRAW_PARAMS = 10 /* .. 10 + 3*NMAX + CMAX */, DRAW_GRAVITY = 40 /*,41*/ };
#define EPISODE_GRAVITY 0xFFFFFFFFu

/* ------------------------------------------------------------------------------------------ */
/* small linear algebra                                                                         */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double m[3][3]; } mat3;
typedef struct { double v[3]; } vec3;

static mat3 m3_mul(mat3 a, mat3 b) {
    mat3 c;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0; for (int k = 0; k < 3; ++k) s += a.m[i][k] * b.m[k][j]; c.m[i][j] = s; }
    return c;
}
static mat3 m3_T(mat3 a) { mat3 c; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) c.m[i][j] = a.m[j][i]; return c; }
static vec3 m3_v(mat3 a, vec3 x) { vec3 y; for (int i = 0; i < 3; ++i) y.v[i] = a.m[i][0]*x.v[0] + a.m[i][1]*x.v[1] + a.m[i][2]*x.v[2]; return y; }
static vec3 v3_cross(vec3 a, vec3 b) { vec3 c = {{a.v[1]*b.v[2]-a.v[2]*b.v[1], a.v[2]*b.v[0]-a.v[0]*b.v[2], a.v[0]*b.v[1]-a.v[1]*b.v[0]}}; return c; }
static vec3 v3_add(vec3 a, vec3 b) { vec3 c = {{a.v[0]+b.v[0], a.v[1]+b.v[1], a.v[2]+b.v[2]}}; return c; }
static vec3 v3_sub(vec3 a, vec3 b) { vec3 c = {{a.v[0]-b.v[0], a.v[1]-b.v[1], a.v[2]-b.v[2]}}; return c; }
static mat3 m3_skew(vec3 a) { mat3 s = {{{0,-a.v[2],a.v[1]},{a.v[2],0,-a.v[0]},{-a.v[1],a.v[0],0}}}; return s; }
static mat3 m3_rot_axis(int axis, double q) {
    double c = cos(q), s = sin(q);
    mat3 r = {{{1,0,0},{0,1,0},{0,0,1}}};
    int a = (axis + 1) % 3, b = (axis + 2) % 3;
    r.m[a][a] = c; r.m[a][b] = -s; r.m[b][a] = s; r.m[b][b] = c;
    return r;
}

/* per-env physical parameters, unpacked from the packed params row */
typedef struct {
    double mass_scale[NMAX], damping[NMAX], friction[NMAX], mu[CMAX], gravity_z;
} env_params;

static void unpack_params(const os2r_model *M, const double *row, env_params *P) {
    int n = M->n_dof, nc = M->n_contacts;
    for (int i = 0; i < n; ++i) { P->mass_scale[i] = row[i]; P->damping[i] = row[n + i]; P->friction[i] = row[2*n + i]; }
    for (int c = 0; c < nc; ++c) P->mu[c] = row[3*n + c];
    P->gravity_z = row[3*n + nc];
}

/* 6x6 spatial inertia of body i about its frame origin, [angular; linear] ordering. Mass is
 * scaled by the randomiser coefficient, the rotational inertia about the COM is NOT
 * (randomizers/monopod.py:183-190 touches link/inertial/mass only). */
static void body_inertia(const os2r_model *M, const env_params *P, int i, double I[6][6]) {
    double m = M->mass[i] * P->mass_scale[i];
    vec3 c = {{M->com[i][0], M->com[i][1], M->com[i][2]}};
    const double *t = M->inertia[i];
    mat3 Ic = {{{t[0], t[3], t[4]}, {t[3], t[1], t[5]}, {t[4], t[5], t[2]}}};
    mat3 cx = m3_skew(c), cxT = m3_T(cx), cc = m3_mul(cx, cxT);
    memset(I, 0, 36 * sizeof(double));
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
        I[a][b] = Ic.m[a][b] + m * cc.m[a][b];
        I[a][3 + b] = m * cx.m[a][b];
        I[3 + a][b] = m * cxT.m[a][b];
    }
    for (int a = 0; a < 3; ++a) I[3 + a][3 + a] = m;
}

/* Pluecker motion transform parent -> child: X = [E 0; -E rx E] */
static void plux(mat3 E, vec3 r, double X[6][6]) {
    mat3 Erx = m3_mul(E, m3_skew(r));
    memset(X, 0, 36 * sizeof(double));
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
        X[a][b] = E.m[a][b]; X[3 + a][3 + b] = E.m[a][b]; X[3 + a][b] = -Erx.m[a][b];
    }
}
static void crm(const double v[6], double C[6][6]) {
    vec3 w = {{v[0], v[1], v[2]}}, l = {{v[3], v[4], v[5]}};
    mat3 wx = m3_skew(w), lx = m3_skew(l);
    memset(C, 0, 36 * sizeof(double));
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
        C[a][b] = wx.m[a][b]; C[3 + a][3 + b] = wx.m[a][b]; C[3 + a][b] = lx.m[a][b];
    }
}
static void mv6(const double A[6][6], const double x[6], double y[6]) {
    for (int a = 0; a < 6; ++a) { double s = 0; for (int b = 0; b < 6; ++b) s += A[a][b] * x[b]; y[a] = s; }
}
static void mTv6(const double A[6][6], const double x[6], double y[6]) {
    for (int a = 0; a < 6; ++a) { double s = 0; for (int b = 0; b < 6; ++b) s += A[b][a] * x[b]; y[a] = s; }
}

/* Articulated-body algorithm on the serial chain (Featherstone, RBDA table 7.1), body coordinates.
 * implicit_dt > 0: DART's implicit joint damping (denominator + dt*d_i, force - d_i*qd_i).
 * with_bias = 0: velocities and gravity ignored (used for the unit-torque Minv columns).       */
static void aba(const os2r_model *M, const env_params *P, const double *q, const double *qd,
                const double *tau, double implicit_dt, int with_bias, double *qdd) {
    int n = M->n_dof;
    double X[NMAX][6][6], IA[NMAX][6][6], pA[NMAX][6], v[NMAX][6], c[NMAX][6], U[NMAX][6], D[NMAX], u[NMAX], a[NMAX][6];
    for (int i = 0; i < n; ++i) {
        mat3 Rt; memcpy(Rt.m, M->tree_R[i], sizeof(Rt.m));
        mat3 E = m3_T(m3_mul(Rt, m3_rot_axis(M->axis[i], q[i])));
        vec3 r = {{M->tree_p[i][0], M->tree_p[i][1], M->tree_p[i][2]}};
        plux(E, r, X[i]);
        double vJ[6] = {0, 0, 0, 0, 0, 0};
        if (with_bias) vJ[M->axis[i]] = qd[i];
        if (i == 0) memcpy(v[i], vJ, sizeof(vJ));
        else { mv6(X[i], v[i - 1], v[i]); for (int k = 0; k < 6; ++k) v[i][k] += vJ[k]; }
        double C[6][6]; crm(v[i], C); mv6(C, vJ, c[i]);
        body_inertia(M, P, i, IA[i]);
        double Iv[6]; mv6(IA[i], v[i], Iv);
        /* pA = v x* (I v) = -crm(v)^T (I v) */
        double t[6]; mTv6(C, Iv, t);
        for (int k = 0; k < 6; ++k) pA[i][k] = with_bias ? -t[k] : 0.0;
    }
    for (int i = n - 1; i >= 0; --i) {
        int ax = M->axis[i];
        for (int k = 0; k < 6; ++k) U[i][k] = IA[i][k][ax];
        double damp = implicit_dt > 0 ? P->damping[i] : 0.0;
        D[i] = U[i][ax] + implicit_dt * damp;
        u[i] = tau[i] - (with_bias ? damp * qd[i] : 0.0) - pA[i][ax];
        if (i > 0) {
            double Ia[6][6], pa[6], t6[6];
            for (int r_ = 0; r_ < 6; ++r_) for (int s = 0; s < 6; ++s) Ia[r_][s] = IA[i][r_][s] - U[i][r_] * U[i][s] / D[i];
            mv6(Ia, c[i], t6);
            for (int k = 0; k < 6; ++k) pa[k] = pA[i][k] + t6[k] + U[i][k] * u[i] / D[i];
            /* IA_parent += X^T Ia X ; pA_parent += X^T pa */
            double T[6][6];
            for (int r_ = 0; r_ < 6; ++r_) for (int s = 0; s < 6; ++s) { double z = 0; for (int k = 0; k < 6; ++k) z += Ia[r_][k] * X[i][k][s]; T[r_][s] = z; }
            for (int r_ = 0; r_ < 6; ++r_) for (int s = 0; s < 6; ++s) { double z = 0; for (int k = 0; k < 6; ++k) z += X[i][k][r_] * T[k][s]; IA[i - 1][r_][s] += z; }
            mTv6(X[i], pa, t6);
            for (int k = 0; k < 6; ++k) pA[i - 1][k] += t6[k];
        }
    }
    for (int i = 0; i < n; ++i) {
        double ap[6];
        if (i == 0) {
            /* a_parent = -a_gravity expressed in world coords: (0,0,-g_z) upward fictitious acceleration */
            double a0[6] = {0, 0, 0, 0, 0, with_bias ? -P->gravity_z : 0.0};
            mv6(X[i], a0, ap);
        } else mv6(X[i], a[i - 1], ap);
        for (int k = 0; k < 6; ++k) ap[k] += c[i][k];
        double s = 0; for (int k = 0; k < 6; ++k) s += U[i][k] * ap[k];
        qdd[i] = (u[i] - s) / D[i];
        memcpy(a[i], ap, sizeof(ap));
        a[i][M->axis[i]] += qdd[i];
    }
}

/* world-frame forward kinematics: rotation / origin of each body frame */
static void fk_world(const os2r_model *M, const double *q, mat3 *Rw, vec3 *pw) {
    mat3 R = {{{1,0,0},{0,1,0},{0,0,1}}};
    vec3 p = {{0, 0, 0}};
    for (int i = 0; i < M->n_dof; ++i) {
        mat3 Rt; memcpy(Rt.m, M->tree_R[i], sizeof(Rt.m));
        vec3 r = {{M->tree_p[i][0], M->tree_p[i][1], M->tree_p[i][2]}};
        p = v3_add(p, m3_v(R, r));
        R = m3_mul(m3_mul(R, Rt), m3_rot_axis(M->axis[i], q[i]));
        Rw[i] = R; pw[i] = p;
    }
}

typedef struct {
    int n_rows;
    int active[RMAX];
    double J[RMAX][NMAX];     /* constraint Jacobian rows                    */
    double MiJt[RMAX][NMAX];  /* Minv * J^T                                  */
    double Arr[RMAX];         /* diagonal of the Delassus matrix             */
    double target[RMAX];      /* desired row velocity                        */
    double cfm[RMAX];
    double depth[CMAX];
} constraint_set;

static void minv_matrix(const os2r_model *M, const env_params *P, const double *q, double Minv[NMAX][NMAX]) {
    int n = M->n_dof;
    for (int j = 0; j < n; ++j) {
        double tau[NMAX] = {0}, col[NMAX];
        tau[j] = 1.0;
        aba(M, P, q, NULL, tau, 0.0, 0, col);
        for (int i = 0; i < n; ++i) Minv[i][j] = col[i];
    }
}

static void build_constraints(const os2r_model *M, const env_params *P, const double *q,
                              double Minv[NMAX][NMAX], constraint_set *S) {
    int n = M->n_dof, nc = M->n_contacts;
    mat3 Rw[NMAX]; vec3 pw[NMAX];
    fk_world(M, q, Rw, pw);
    memset(S, 0, sizeof(*S));
    S->n_rows = n + 3 * nc;
    for (int i = 0; i < n; ++i) {           /* Coulomb joint friction rows */
        S->active[i] = P->friction[i] > 0.0;
        S->J[i][i] = 1.0;
        S->cfm[i] = M->cfm_joint;
    }
    for (int c = 0; c < nc; ++c) {
        int b = M->contact_body[c];
        vec3 lc = {{M->contact_pos[c][0], M->contact_pos[c][1], M->contact_pos[c][2]}};
        vec3 centre = v3_add(pw[b], m3_v(Rw[b], lc));
        double depth = M->contact_radius[c] - centre.v[2];
        S->depth[c] = depth;
        int act = depth > 0.0;
        vec3 x = centre; x.v[2] -= M->contact_radius[c];   /* lowest point of the sphere */
        for (int k = 0; k < 3; ++k) { S->active[n + 3*c + k] = act; S->cfm[n + 3*c + k] = M->cfm_contact; }
        if (!act) continue;
        for (int i = 0; i <= b; ++i) {
            vec3 ax = {{Rw[i].m[0][M->axis[i]], Rw[i].m[1][M->axis[i]], Rw[i].m[2][M->axis[i]]}};
            vec3 jc = v3_cross(ax, v3_sub(x, pw[i]));  /* velocity of x per unit joint rate */
            S->J[n + 3*c + 0][i] = jc.v[2];   /* normal  = world z */
            S->J[n + 3*c + 1][i] = jc.v[0];   /* tangent = world x */
            S->J[n + 3*c + 2][i] = jc.v[1];   /* tangent = world y */
        }
        double bounce = depth * M->erp / M->dt;
        if (bounce > M->max_erv) bounce = M->max_erv;
        S->target[n + 3*c] = bounce;
    }
    for (int r = 0; r < S->n_rows; ++r) {
        if (!S->active[r]) continue;
        double d = 0;
        for (int i = 0; i < n; ++i) { double s = 0; for (int j = 0; j < n; ++j) s += Minv[i][j] * S->J[r][j]; S->MiJt[r][i] = s; }
        for (int i = 0; i < n; ++i) d += S->J[r][i] * S->MiJt[r][i];
        S->Arr[r] = d;
    }
}

/* Diagnostic: how many sweeps each physics iteration ran (index = sweeps, last bin = overflow). */
static long long g_sweep_hist[65];
void oracle_sweep_histogram(long long out[65], int clear) {
    for (int i = 0; i < 65; ++i) { out[i] = __atomic_load_n(&g_sweep_hist[i], __ATOMIC_RELAXED); if (clear) __atomic_store_n(&g_sweep_hist[i], 0, __ATOMIC_RELAXED); }
}

/* Projected Gauss-Seidel, fixed row order, at most `sweeps` sweeps. The iteration of this env ends after
 * the first sweep whose velocity change is <= tol in the kinetic-energy norm sqrt(dv^T M dv) (the kernel
 * measures the same quantity as |dz| in its Cholesky-whitened coordinates); tol = 0 ends it only when a
 * sweep left the velocity exactly unchanged (os2r_model.pgs_tol, include/os2r.h). The norm needs no mass
 * matrix: dv = Minv J^T dlam, so dv^T M dv = sum_r dlam_r (J_r . dv). */
static void pgs_sweeps(const os2r_model *M, const env_params *P, const constraint_set *S,
                       double Minv[NMAX][NMAX], double *v, double *lam, int sweeps, double tol) {
    int n = M->n_dof, it;
    (void)Minv;
    for (it = 0; it < sweeps; ) {
        double v0[NMAX], dlam[RMAX] = {0};
        for (int i = 0; i < n; ++i) v0[i] = v[i];
        /* the joint-friction rows take part in the first pgs_joint_sweeps sweeps only (os2r_model.pgs_joint_sweeps) */
        for (int r = (M->pgs_joint_sweeps > 0 && it >= M->pgs_joint_sweeps) ? n : 0; r < S->n_rows; ++r) {
            if (!S->active[r]) continue;
            double lo, hi;
            if (r < n) { hi = P->friction[r] * M->dt; lo = -hi; }
            else {
                int c = (r - n) / 3, k = (r - n) % 3;
                if (k == 0) { lo = 0; hi = INFINITY; }
                else { hi = P->mu[c] * lam[n + 3*c]; lo = -hi; }
            }
            double w = -S->target[r] + S->cfm[r] * S->Arr[r] * lam[r];
            for (int i = 0; i < n; ++i) w += S->J[r][i] * v[i];
            double nl = lam[r] - w / (S->Arr[r] * (1.0 + S->cfm[r]));
            if (nl < lo) nl = lo;
            if (nl > hi) nl = hi;
            double dl = nl - lam[r];
            for (int i = 0; i < n; ++i) v[i] += S->MiJt[r][i] * dl;
            lam[r] = nl;
            dlam[r] = dl;
        }
        ++it;
        double e2 = 0;
        for (int r = 0; r < S->n_rows; ++r) {
            if (dlam[r] == 0.0) continue;
            double jdv = 0;
            for (int i = 0; i < n; ++i) jdv += S->J[r][i] * (v[i] - v0[i]);
            e2 += dlam[r] * jdv;
        }
        if (e2 <= tol * tol) break;
    }
    __atomic_fetch_add(&g_sweep_hist[it < 64 ? it : 64], 1, __ATOMIC_RELAXED);
}

/* One physics iteration (dt). state: q, v, lam (warm start) updated in place. */
static void substep(const os2r_model *M, const env_params *P, double *q, double *v, double *lam,
                    const double *action, int sweeps, double tol) {
    int n = M->n_dof;
    double tau[NMAX] = {0}, qdd[NMAX], Minv[NMAX][NMAX];
    if (M->role_dof[OS2R_ROLE_HIP] >= 0) tau[M->role_dof[OS2R_ROLE_HIP]] = M->max_torque[0] * action[0];
    if (M->role_dof[OS2R_ROLE_KNEE] >= 0) tau[M->role_dof[OS2R_ROLE_KNEE]] = M->max_torque[1] * action[1];
    aba(M, P, q, v, tau, M->dt, 1, qdd);
    for (int i = 0; i < n; ++i) v[i] += M->dt * qdd[i];
    minv_matrix(M, P, q, Minv);
    constraint_set S;
    build_constraints(M, P, q, Minv, &S);
    for (int r = 0; r < S.n_rows; ++r) {
        if (!S.active[r]) { lam[r] = 0.0; continue; }
        for (int i = 0; i < n; ++i) v[i] += S.MiJt[r][i] * lam[r];   /* warm start */
    }
    pgs_sweeps(M, P, &S, Minv, v, lam, sweeps, tol);
    for (int i = 0; i < n; ++i) q[i] += M->dt * v[i];
}

/* ------------------------------------------------------------------------------------------ */
/* task: observation / reward / done  (tasks/monopod.py, rewards/)                              */
/* ------------------------------------------------------------------------------------------ */
static const double PI = 3.141592653589793;  /* np.pi */
static const double EPS = 2.220446049250313e-16; /* np.finfo(float).eps */

/* np.mod(x, y) for y > 0 */
static double np_mod(double x, double y) {
    double r = fmod(x, y);
    if (r != 0.0) { if (r < 0.0) r += y; } else r = 0.0;
    return r;
}

/* raw (masked, wrapped) and final observation; tasks/monopod.py:248-271, monopod_no_norm.py:222-246 */
static void observe(const os2r_model *M, const os2r_task_cfg *T, const double *q, const double *v,
                    const double *a_old, double *raw, double *obs) {
    (void)M;
    for (int k = 0; k < T->obs_dim; ++k) {
        double x;
        switch (T->obs_kind[k]) {
        case OS2R_OBS_POS: x = q[T->obs_index[k]]; break;
        case OS2R_OBS_POS_PERIODIC: x = np_mod(q[T->obs_index[k]] + PI, 2 * PI) - PI; break;
        case OS2R_OBS_VEL: x = v[T->obs_index[k]]; break;
        default: x = a_old[T->obs_index[k]]; break;
        }
        raw[k] = x;
        if (!T->normalized) obs[k] = x;
        else if (T->obs_kind[k] == OS2R_OBS_VEL) obs[k] = tanh(0.05 * x);
        else obs[k] = 2 * (x - T->obs_low[k]) / (T->obs_high[k] - T->obs_low[k]) - 1;
    }
}

/* done = not reset_space.contains(obs); reset_space = Box(low+eps, high-eps) (monopod.py:198,284-286) */
static int is_done(const os2r_task_cfg *T, const double *obs) {
    for (int k = 0; k < T->obs_dim; ++k) {
        double lo, hi;
        if (T->normalized) { lo = -1.0 + EPS; hi = 1.0 - EPS; }
        else if (T->obs_kind[k] == OS2R_OBS_VEL) { lo = -INFINITY; hi = INFINITY; }
        else { lo = T->obs_low[k] + EPS; hi = T->obs_high[k] - EPS; }
        if (!(obs[k] >= lo) || !(obs[k] <= hi)) return 1;
    }
    return 0;
}

enum { SIG_GAUSSIAN = 0, SIG_HYPERBOLIC, SIG_LONG_TAIL, SIG_RECIPROCAL, SIG_COSINE, SIG_LINEAR, SIG_QUADRATIC, SIG_TANH_SQUARED };

/* rewards/rewards_utils.py:10-73 */
static double sigmoid_fn(double x, double value_at_1, int kind) {
    double scale, sx;
    switch (kind) {
    case SIG_GAUSSIAN: scale = sqrt(-2 * log(value_at_1)); return exp(-0.5 * (x * scale) * (x * scale));
    case SIG_HYPERBOLIC: scale = acosh(1 / value_at_1); return 1 / cosh(x * scale);
    case SIG_LONG_TAIL: scale = sqrt(1 / value_at_1 - 1); return 1 / ((x * scale) * (x * scale) + 1);
    case SIG_RECIPROCAL: scale = 1 / value_at_1 - 1; return 1 / (fabs(x) * scale + 1);
    case SIG_COSINE: scale = acos(2 * value_at_1 - 1) / PI; sx = x * scale; return fabs(sx) < 1 ? (1 + cos(PI * sx)) / 2 : 0.0;
    case SIG_LINEAR: scale = 1 - value_at_1; sx = x * scale; return fabs(sx) < 1 ? 1 - sx : 0.0;
    case SIG_QUADRATIC: scale = sqrt(1 - value_at_1); sx = x * scale; return fabs(sx) < 1 ? 1 - sx * sx : 0.0;
    default: scale = atanh(sqrt(1 - value_at_1)); { double t = tanh(x * scale); return 1 - t * t; }
    }
}
/* rewards/rewards_utils.py:76-122 */
double oracle_tolerance(double x, double lower, double upper, double margin, int sigmoid, double value_at_margin) {
    int in_bounds = (lower <= x) && (x <= upper);
    if (margin == 0) return in_bounds ? 1.0 : 0.0;
    double d = (x < lower ? lower - x : x - upper) / margin;
    return in_bounds ? 1.0 : sigmoid_fn(d, value_at_margin, sigmoid);
}

/* rewards/__init__.py:66-207; a0 = actions[0] (current), a1 = actions[1] (previous) */
static double reward_fn(const os2r_task_cfg *T, const double *obs, const double *a0, const double *a1) {
    double H = T->normalized ? 0.11 / 1.57 : 0.11;
    double bp = T->reward_pitch_col >= 0 ? obs[T->reward_pitch_col] : 0.0;
    switch (T->reward_id) {
    case OS2R_REWARD_BALANCING_V1:
        return oracle_tolerance(bp, H, 4 * H, 0, SIG_GAUSSIAN, 0.1);
    case OS2R_REWARD_BALANCING_V2: {
        double r = oracle_tolerance(bp, H, 4 * H, 0, SIG_GAUSSIAN, 0.1);
        for (int j = 0; j < 2; ++j) r *= oracle_tolerance(a0[j], 0, 0, 1, SIG_QUADRATIC, 0.4);
        return r; }
    case OS2R_REWARD_BALANCING_V3: {
        double r = oracle_tolerance(bp, H, 4 * H, 0.01, SIG_LONG_TAIL, 0.1);
        for (int j = 0; j < 2; ++j) r *= oracle_tolerance(a0[j] - a1[j], 0, 0, 1, SIG_QUADRATIC, 0.1);
        return r; }
    case OS2R_REWARD_HOPPING_V1: {
        double r = oracle_tolerance(bp, H, 4 * H, 0, SIG_GAUSSIAN, 0.1);
        for (int j = 0; j < 2; ++j) r *= oracle_tolerance(a0[j] - a1[j], 0, 0, 0.1, SIG_QUADRATIC, 0.0);
        r *= oracle_tolerance(obs[T->reward_yawvel_col], 0.25, 0.3, 0.15, SIG_TANH_SQUARED, 0.1);
        return r; }
    case OS2R_REWARD_STRAIGHT_V1: {
        double sc = 0;
        for (int j = 0; j < 2; ++j) sc += oracle_tolerance(a0[j] / 20, 0, 0, 1, SIG_QUADRATIC, 0.0);
        sc = (4 + sc / 2) / 5;
        double hr = oracle_tolerance(obs[T->reward_hip_col], 0, 0, 1, SIG_LINEAR, 0.1);
        double kr = oracle_tolerance(obs[T->reward_knee_col], 0, 0, 1, SIG_LINEAR, 0.1);
        return hr * kr * sc; }
    default: return 0.0;
    }
}

/* utils/reset.py:4-40 (arguments of acos/asin clamped to [-1,1]: unreachable with the shipped resets) */
void oracle_leg_joint_angles(const os2r_task_cfg *T, double bp, double out[2]) {
    double lh = (T->ik_boom * sin(bp) + T->ik_pivot_height) / cos(bp);
    double ul = T->ik_upper_leg, ll = T->ik_lower_leg;
    double lleg = lh - T->ik_hip_offset - T->ik_clip;
    if (lleg > ul + ll) { out[0] = 0; out[1] = 0; return; }
    double ca = (ul * ul + lleg * lleg - ll * ll) / (2 * ul * lleg);
    if (ca > 1) ca = 1; if (ca < -1) ca = -1;
    double hip = acos(ca);
    double sa = ul * sin(hip) / ll;
    if (sa > 1) sa = 1; if (sa < -1) sa = -1;
    out[0] = hip;
    out[1] = -(asin(sa) + hip);
}

/* ------------------------------------------------------------------------------------------ */
/* reset + randomisation                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {            /* per-env episode bookkeeping (one row each, arrays owned by caller) */
    int32_t *steps;         /* steps in the current episode                                      */
    double *ret;            /* return of the current episode                                     */
    uint32_t *episode;      /* episode counter = RNG counter word                                */
    int32_t *reset_id;      /* index into the task's reset_positions                             */
} episode_arrays;

static void draw_params(const os2r_model *M, const os2r_task_cfg *T, uint64_t seed, uint64_t gid,
                        uint32_t episode, double *prow) {
    int n = M->n_dof, nc = M->n_contacts;
    /* randomizers/monopod.py:182-215 — mass coef, friction absolute, damping coef (zeros skipped), mu coef */
    for (int i = 0; i < n; ++i) {
        if (T->randomize_params) {
            double um = rng_uniform(seed, gid, episode, DRAW_PARAMS + i);
            double uf = rng_uniform(seed, gid, episode, DRAW_PARAMS + NMAX + i);
            double ud = rng_uniform(seed, gid, episode, DRAW_PARAMS + 2 * NMAX + i);
            prow[i] = T->mass_lo + (T->mass_hi - T->mass_lo) * um;
            prow[2*n + i] = T->fric_lo + (T->fric_hi - T->fric_lo) * uf;
            prow[n + i] = M->damping[i] * (T->damp_lo + (T->damp_hi - T->damp_lo) * ud);
        } else {
            prow[i] = 1.0; prow[n + i] = M->damping[i]; prow[2*n + i] = M->friction[i];
        }
    }
    for (int c = 0; c < nc; ++c) {
        if (T->randomize_params) {
            double uc = rng_uniform(seed, gid, episode, DRAW_PARAMS + 3 * NMAX + c);
            double mu_l = T->mu_link * (T->mu_lo + (T->mu_hi - T->mu_lo) * uc);
            prow[3*n + c] = mu_l;   /* min(link, ground=1): link value (< 1) wins */
        } else prow[3*n + c] = M->contact_mu[c];
    }
}

static void draw_gravity(const os2r_model *M, const os2r_task_cfg *T, uint64_t seed, uint64_t gid, double *prow) {
    int n = M->n_dof, nc = M->n_contacts;
    if (T->randomize_gravity) {   /* randomizers/monopod.py:56-61 */
        double z[2]; rng_normal2(seed, gid, EPISODE_GRAVITY, 0, z);
        prow[3*n + nc] = T->grav_mean + T->grav_std * z[0];
    } else prow[3*n + nc] = M->gravity_z;
}

static void reset_env(const os2r_model *M, const os2r_task_cfg *T, uint64_t seed, uint64_t gid,
                      double *srow, double *prow, const episode_arrays *E, int64_t e) {
    int n = M->n_dof, rows = n + 3 * M->n_contacts;
    uint32_t ep = ++E->episode[e];
    double *q = srow, *v = srow + n, *lam = srow + 2*n;
    for (int i = 0; i < n; ++i) { q[i] = 0; v[i] = 0; }
    for (int r = 0; r < rows; ++r) lam[r] = 0;
    /* a_prev (action history) is NOT cleared: the reference keeps the deque across resets */
    int idx = (int)(rng_uniform(seed, gid, ep, DRAW_RESET) * T->n_resets);
    if (idx >= T->n_resets) idx = T->n_resets - 1;
    E->reset_id[e] = idx;
    E->steps[e] = 0; E->ret[e] = 0;
    double pitch = T->reset_pitch[idx];
    double leg[2];
    if (T->reset_randomized) {
        pitch *= 0.8 + 0.4 * rng_uniform(seed, gid, ep, DRAW_PITCH);
        double z[2]; rng_normal2(seed, gid, ep, DRAW_NOISE, z);
        double r0 = fabs(0.2 * z[0]), r1 = fabs(0.2 * z[1]);
        double rmax = r0 > r1 ? r0 : r1, rmin = r0 > r1 ? r1 : r0;
        if (!T->reset_laying[idx]) oracle_leg_joint_angles(T, pitch, leg);
        else { leg[0] = 1.57 - (rng_uniform(seed, gid, ep, DRAW_LAYSIDE) < 0.5 ? 3.14 : 0.0); leg[1] = 0; }
        /* precedence quirk of randomizers/monopod.py:102-103: (a>0 - a<0) == (a>0) */
        leg[0] = leg[0] + (leg[0] > 0 ? 1.0 : 0.0) * rmax;
        leg[1] = leg[1] - (leg[1] > 0 ? 1.0 : 0.0) * rmin;
        double dir = 1.0 - (rng_uniform(seed, gid, ep, DRAW_DIR) < 0.5 ? 2.0 : 0.0);
        leg[0] *= dir; leg[1] *= dir;
        if (M->role_dof[OS2R_ROLE_YAW] >= 0) q[M->role_dof[OS2R_ROLE_YAW]] = -0.2 + 0.4 * rng_uniform(seed, gid, ep, DRAW_YAW);
    } else if (T->simple_sample_reset) {
        leg[0] = T->simple_lo[0] + (T->simple_hi[0] - T->simple_lo[0]) * rng_uniform(seed, gid, ep, DRAW_SIMPLE_HIP);
        leg[1] = T->simple_lo[1] + (T->simple_hi[1] - T->simple_lo[1]) * rng_uniform(seed, gid, ep, DRAW_SIMPLE_KNEE);
    } else {
        if (!T->reset_laying[idx]) oracle_leg_joint_angles(T, pitch, leg);
        else { leg[0] = 1.57; leg[1] = 0; }
    }
    if (M->role_dof[OS2R_ROLE_PITCH] >= 0) q[M->role_dof[OS2R_ROLE_PITCH]] = pitch;
    if (M->role_dof[OS2R_ROLE_HIP] >= 0) q[M->role_dof[OS2R_ROLE_HIP]] = leg[0];
    if (M->role_dof[OS2R_ROLE_KNEE] >= 0) q[M->role_dof[OS2R_ROLE_KNEE]] = leg[1];
    draw_params(M, T, seed, gid, ep, prow);
    /* MonopodEnvRandomizer(num_physics_rollouts=K): randomize_physics again at every K-th reset
     * (randomizers/monopod.py:36,56-61,371; gym-ignition's physics_expired counter) */
    if (T->randomize_gravity && T->gravity_redraw_resets > 0 && ep % (uint32_t)T->gravity_redraw_resets == 0) {
        double z[2]; rng_normal2(seed, gid, ep, DRAW_GRAVITY, z);
        prow[3*n + M->n_contacts] = T->grav_mean + T->grav_std * z[0];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* exported batch entry points (ctypes)                                                         */
/* ------------------------------------------------------------------------------------------ */
int oracle_state_width(const os2r_model *M) { return 2 * M->n_dof + (M->n_dof + 3 * M->n_contacts) + 2; }
int oracle_params_width(const os2r_model *M) { return 3 * M->n_dof + M->n_contacts + 1; }

/* initial parameters: nominal + gravity draw (what GazeboEnvRandomizer.__init__ does once) */
void oracle_init(const os2r_model *M, const os2r_task_cfg *T, int64_t N, int64_t first_env_id, uint64_t seed,
                 double *state, double *params, int32_t *steps, double *ret, uint32_t *episode, int32_t *reset_id) {
    int W = oracle_state_width(M), PW = oracle_params_width(M);
    os2r_task_cfg nominal = *T; nominal.randomize_params = 0;
    for (int64_t e = 0; e < N; ++e) {
        memset(state + e * W, 0, W * sizeof(double));
        draw_params(M, &nominal, seed, (uint64_t)(first_env_id + e), 0, params + e * PW);
        draw_gravity(M, T, seed, (uint64_t)(first_env_id + e), params + e * PW);
        steps[e] = 0; ret[e] = 0; episode[e] = 0; reset_id[e] = 0;
    }
}

void oracle_reset(const os2r_model *M, const os2r_task_cfg *T, int64_t N, int64_t first_env_id, uint64_t seed,
                  double *state, double *params, int32_t *steps, double *ret, uint32_t *episode, int32_t *reset_id,
                  const uint8_t *mask, double *obs) {
    int W = oracle_state_width(M), PW = oracle_params_width(M), n = M->n_dof;
    episode_arrays E = {steps, ret, episode, reset_id};
    for (int64_t e = 0; e < N; ++e) {
        if (mask && !mask[e]) continue;
        double *srow = state + e * W;
        reset_env(M, T, seed, (uint64_t)(first_env_id + e), srow, params + e * PW, &E, e);
        if (obs) { double raw[OS2R_MAX_OBS]; observe(M, T, srow, srow + n, srow + 2*n + (n + 3*M->n_contacts), raw, obs + e * T->obs_dim); }
    }
}

/* One env step for N envs (GazeboRuntime.step + SubprocVecEnv auto-reset). info[e] = {reset id, cause bits}. */
typedef struct {
    const os2r_model *M; const os2r_task_cfg *T; int64_t lo, hi, first_env_id; uint64_t seed;
    double *state, *params; int32_t *steps; double *ret; uint32_t *episode; int32_t *reset_id;
    const double *actions; double *obs, *reward; uint8_t *done; double *terminal_obs; int32_t *info;
} step_job;

static void *step_range(void *arg) {
    step_job *J = (step_job *)arg;
    const os2r_model *M = J->M; const os2r_task_cfg *T = J->T;
    int W = oracle_state_width(M), PW = oracle_params_width(M), n = M->n_dof, rows = n + 3 * M->n_contacts, D = T->obs_dim;
    episode_arrays E = {J->steps, J->ret, J->episode, J->reset_id};
    for (int64_t e = J->lo; e < J->hi; ++e) {
        double *srow = J->state + e * W, *q = srow, *v = srow + n, *lam = srow + 2*n, *a_prev = srow + 2*n + rows;
        env_params P; unpack_params(M, J->params + e * PW, &P);
        double a[2] = {J->actions[2*e], J->actions[2*e + 1]}, a_old[2] = {a_prev[0], a_prev[1]};
        /* a non-finite action (rejected by the reference's assert, tasks/monopod.py:218) applies no torque and takes
         * the non-finite path: zero observation / reward, forced reset; valid ones are clipped to the force limits */
        int finite = isfinite(a[0]) && isfinite(a[1]);
        if (!finite) { a[0] = 0; a[1] = 0; }
        for (int k = 0; k < 2; ++k) { if (a[k] > 1) a[k] = 1; if (a[k] < -1) a[k] = -1; }
        for (int s = 0; s < M->substeps; ++s) substep(M, &P, q, v, lam, a, M->pgs_iters, M->pgs_tol);
        for (int i = 0; i < n; ++i) if (!isfinite(q[i]) || !isfinite(v[i])) finite = 0;
        if (!finite) for (int i = 0; i < n; ++i) { q[i] = 0; v[i] = 0; }
        double raw[OS2R_MAX_OBS], o[OS2R_MAX_OBS];
        observe(M, T, q, v, a_old, raw, o);
        double r = finite ? reward_fn(T, o, a, a_old) : 0.0;
        int cause = (finite && is_done(T, o)) ? 1 : 0;
        if (!finite) cause = 4;
        J->steps[e] += 1; J->ret[e] += r;
        if (T->max_episode_steps > 0 && J->steps[e] >= T->max_episode_steps) cause |= 2;
        a_prev[0] = a[0]; a_prev[1] = a[1];
        J->reward[e] = r; J->done[e] = cause != 0;
        if (J->terminal_obs) memcpy(J->terminal_obs + e * D, o, D * sizeof(double));
        if (cause && (T->auto_reset || !finite)) {
            reset_env(M, T, J->seed, (uint64_t)(J->first_env_id + e), srow, J->params + e * PW, &E, e);
            observe(M, T, q, v, a_old, raw, o);
        }
        memcpy(J->obs + e * D, o, D * sizeof(double));
        if (J->info) { J->info[2*e] = J->reset_id[e]; J->info[2*e + 1] = cause; }
    }
    return NULL;
}

void oracle_step(const os2r_model *M, const os2r_task_cfg *T, int64_t N, int64_t first_env_id, uint64_t seed,
                 double *state, double *params, int32_t *steps, double *ret, uint32_t *episode, int32_t *reset_id,
                 const double *actions, double *obs, double *reward, uint8_t *done, double *terminal_obs,
                 int32_t *info, int32_t nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > N) nthreads = N > 0 ? (int32_t)N : 1;
    step_job jobs[256]; pthread_t th[256];
    for (int t = 0; t < nthreads; ++t) {
        step_job j = {M, T, N * t / nthreads, N * (t + 1) / nthreads, first_env_id, seed, state, params, steps, ret,
                      episode, reset_id, actions, obs, reward, done, terminal_obs, info};
        jobs[t] = j;
    }
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, step_range, &jobs[t]);
    step_range(&jobs[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
}

/* ---- unit-test hooks ---------------------------------------------------------------------- */
/* forward dynamics pieces at one state: qdd (implicit damping), Minv, contact depths, J rows */
void oracle_dynamics_debug(const os2r_model *M, const double *prow, const double *q, const double *v,
                           const double *action, double *qdd, double *Minv_out, double *depth, double *J_out) {
    env_params P; unpack_params(M, prow, &P);
    int n = M->n_dof;
    double tau[NMAX] = {0}, Minv[NMAX][NMAX];
    if (M->role_dof[OS2R_ROLE_HIP] >= 0) tau[M->role_dof[OS2R_ROLE_HIP]] = M->max_torque[0] * action[0];
    if (M->role_dof[OS2R_ROLE_KNEE] >= 0) tau[M->role_dof[OS2R_ROLE_KNEE]] = M->max_torque[1] * action[1];
    aba(M, &P, q, v, tau, M->dt, 1, qdd);
    minv_matrix(M, &P, q, Minv);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Minv_out[i * n + j] = Minv[i][j];
    constraint_set S; build_constraints(M, &P, q, Minv, &S);
    for (int c = 0; c < M->n_contacts; ++c) depth[c] = S.depth[c];
    for (int r = 0; r < S.n_rows; ++r) for (int i = 0; i < n; ++i) J_out[r * n + i] = S.J[r][i];
}
/* forward dynamics without damping/implicit terms: qdd = M^-1 (tau - h); for energy tests */
void oracle_forward_dynamics_plain(const os2r_model *M, const double *prow, const double *q, const double *v,
                                   const double *tau, double *qdd) {
    env_params P; unpack_params(M, prow, &P);
    aba(M, &P, q, v, tau, 0.0, 1, qdd);
}
/* n physics iterations on one env with a constant action; sweeps<=0 uses model->pgs_iters; tol>0 = converge */
void oracle_substeps(const os2r_model *M, const double *prow, double *q, double *v, double *lam,
                     const double *action, int count, int sweeps, double tol) {
    env_params P; unpack_params(M, prow, &P);
    for (int s = 0; s < count; ++s) substep(M, &P, q, v, lam, action, sweeps > 0 ? sweeps : M->pgs_iters, tol);
}
/* observation / reward / done of a given state (no physics) */
void oracle_evaluate(const os2r_model *M, const os2r_task_cfg *T, const double *q, const double *v,
                     const double *a0, const double *a1, double *raw, double *obs, double *reward, int32_t *done) {
    observe(M, T, q, v, a1, raw, obs);
    *reward = reward_fn(T, obs, a0, a1);
    *done = is_done(T, obs);
}
/* reward / done of a given observation vector (task.get_state_info, tasks/monopod.py:348-366) */
void oracle_state_info(const os2r_task_cfg *T, const double *obs, const double *a0, const double *a1,
                       double *reward, int32_t *done) {
    *reward = reward_fn(T, obs, a0, a1);
    *done = is_done(T, obs);
}
/* world position of every body origin and contact-sphere centre */
void oracle_fk(const os2r_model *M, const double *q, double *body_pos, double *contact_centre) {
    mat3 Rw[NMAX]; vec3 pw[NMAX];
    fk_world(M, q, Rw, pw);
    for (int i = 0; i < M->n_dof; ++i) for (int k = 0; k < 3; ++k) body_pos[3*i + k] = pw[i].v[k];
    for (int c = 0; c < M->n_contacts; ++c) {
        int b = M->contact_body[c];
        vec3 lc = {{M->contact_pos[c][0], M->contact_pos[c][1], M->contact_pos[c][2]}};
        vec3 x = v3_add(pw[b], m3_v(Rw[b], lc));
        for (int k = 0; k < 3; ++k) contact_centre[3*c + k] = x.v[k];
    }
}
/* total mechanical energy (kinetic + potential) for conservation tests */
double oracle_energy(const os2r_model *M, const double *prow, const double *q, const double *v) {
    env_params P; unpack_params(M, prow, &P);
    int n = M->n_dof;
    mat3 Rw[NMAX]; vec3 pw[NMAX];
    fk_world(M, q, Rw, pw);
    /* body twists in world axes about world origin */
    double E = 0;
    vec3 w = {{0,0,0}}, vo = {{0,0,0}};   /* angular velocity, velocity of the point at world origin */
    for (int i = 0; i < n; ++i) {
        vec3 ax = {{Rw[i].m[0][M->axis[i]], Rw[i].m[1][M->axis[i]], Rw[i].m[2][M->axis[i]]}};
        for (int k = 0; k < 3; ++k) w.v[k] += ax.v[k] * v[i];
        vec3 lin = v3_cross(pw[i], ax);  /* p x a = velocity at origin per unit rate */
        for (int k = 0; k < 3; ++k) vo.v[k] += lin.v[k] * v[i];
        double m = M->mass[i] * P.mass_scale[i];
        vec3 lc = {{M->com[i][0], M->com[i][1], M->com[i][2]}};
        vec3 c = v3_add(pw[i], m3_v(Rw[i], lc));
        vec3 vc = v3_add(vo, v3_cross(w, c));
        const double *t = M->inertia[i];
        mat3 Ic = {{{t[0], t[3], t[4]}, {t[3], t[1], t[5]}, {t[4], t[5], t[2]}}};
        vec3 wb = m3_v(m3_T(Rw[i]), w);
        vec3 Iw = m3_v(Ic, wb);
        E += 0.5 * m * (vc.v[0]*vc.v[0] + vc.v[1]*vc.v[1] + vc.v[2]*vc.v[2]);
        E += 0.5 * (wb.v[0]*Iw.v[0] + wb.v[1]*Iw.v[1] + wb.v[2]*Iw.v[2]);
        E += -m * P.gravity_z * c.v[2];
    }
    return E;
}
