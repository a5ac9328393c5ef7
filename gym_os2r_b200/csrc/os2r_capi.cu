// os2r_capi.cu — C-ABI (include/os2r.h) over the CUDA kernels. Plain pointers and sizes only.
//
// Ownership: the handle owns the SoA env state / parameters / statistics in HBM plus a pinned
// staging area for os2r_step_host; callers own every I/O buffer. There is NO CPU fallback: a
// missing device or a failed launch is reported through the status code + os2r_last_error().
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "os2r_kernels.h"

using namespace os2r;

namespace {

thread_local char g_err[512] = "";

int fail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

#define CK(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// rotation taking child-frame x to the original joint axis (x: identity, y, z: proper rotations)
void axis_normaliser(int axis, double C[9]) {
    static const double Cx[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    static const double Cy[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};   // C e_x = e_y
    static const double Cz[9] = {0, 0, -1, 0, 1, 0, 1, 0, 0};   // C e_x = e_z
    memcpy(C, axis == 0 ? Cx : (axis == 1 ? Cy : Cz), 9 * sizeof(double));
}
void mat_mul(const double *A, const double *B, double *O) {   // 3x3 row-major
    double t[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t[3*r+c] = A[3*r]*B[c] + A[3*r+1]*B[3+c] + A[3*r+2]*B[6+c];
    memcpy(O, t, sizeof(t));
}
void mat_T(const double *A, double *O) { double t[9]; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t[3*r+c] = A[3*c+r]; memcpy(O, t, sizeof(t)); }
void mat_vec(const double *A, const double *x, double *y) { double t[3]; for (int r = 0; r < 3; ++r) t[r] = A[3*r]*x[0] + A[3*r+1]*x[1] + A[3*r+2]*x[2]; memcpy(y, t, sizeof(t)); }

// Host-side normalisation: re-express every body frame so that its joint rotates about child x.
// New body frame B' = B*C (C e_x = axis): R_tree' = Cprev^T R_tree C, p_tree' = Cprev^T p_tree,
// com' = C^T com, I' = C^T I C, contact' = C^T contact.
template <typename T>
void build_model_dev(const os2r_model &m, const os2r_tuning &tune, ModelDev<T> &d) {
    memset(&d, 0, sizeof(d));
    double Cprev[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double Cs[OS2R_MAX_DOF][9];
    for (int i = 0; i < m.n_dof; ++i) {
        double C[9], CprevT[9], CT[9], R[9], p[3], com[3];
        axis_normaliser(m.axis[i], C);
        memcpy(Cs[i], C, sizeof(C));
        mat_T(Cprev, CprevT);
        mat_T(C, CT);
        mat_mul(CprevT, m.tree_R[i], R);
        mat_mul(R, C, R);
        mat_vec(CprevT, m.tree_p[i], p);
        mat_vec(CT, m.com[i], com);
        const double *t = m.inertia[i];
        double I[9] = {t[0], t[3], t[4], t[3], t[1], t[5], t[4], t[5], t[2]}, I2[9];
        mat_mul(CT, I, I2);
        mat_mul(I2, C, I2);
        for (int k = 0; k < 9; ++k) d.tree_R[i][k] = (T)R[k];
        for (int k = 0; k < 3; ++k) { d.tree_p[i][k] = (T)p[k]; d.com[i][k] = (T)com[k]; }
        d.inertia[i][0] = (T)I2[0]; d.inertia[i][1] = (T)I2[4]; d.inertia[i][2] = (T)I2[8];
        d.inertia[i][3] = (T)I2[1]; d.inertia[i][4] = (T)I2[2]; d.inertia[i][5] = (T)I2[5];
        d.mass[i] = (T)m.mass[i];
        memcpy(Cprev, C, sizeof(C));
    }
    for (int c = 0; c < m.n_contacts; ++c) {
        double CT[9], cp[3];
        mat_T(Cs[m.contact_body[c]], CT);
        mat_vec(CT, m.contact_pos[c], cp);
        for (int k = 0; k < 3; ++k) d.contact_pos[c][k] = (T)cp[k];
        d.contact_radius[c] = (T)m.contact_radius[c];
        d.contact_body[c] = m.contact_body[c];
    }
    d.dt = (T)m.dt;
    d.erp_over_dt = (T)(m.erp / m.dt);
    d.max_erv = (T)m.max_erv;
    d.cfm_contact = (T)m.cfm_contact;
    d.cfm_joint = (T)m.cfm_joint;
    d.cfm1_contact = (T)(1.0 + m.cfm_contact); d.cfm1_joint = (T)(1.0 + m.cfm_joint);
    d.kc1 = (T)1 / d.cfm1_contact; d.kj1 = (T)1 / d.cfm1_joint;    // same rounding as the division the kernel used to do
    d.max_torque[0] = (T)m.max_torque[0];
    d.max_torque[1] = (T)m.max_torque[1];
    d.hip_dof = m.role_dof[OS2R_ROLE_HIP];
    d.knee_dof = m.role_dof[OS2R_ROLE_KNEE];
    d.substeps = m.substeps;
    d.pgs_iters = m.pgs_iters;
    d.pgs_joint_sweeps = m.pgs_joint_sweeps;
    d.pgs_tol2 = (T)(m.pgs_tol * m.pgs_tol);
    // lane-sorting hint: a proxy within this clearance may touch down during the next env step
    d.sort_margin = (T)(tune.sort_margin > 0.0 ? tune.sort_margin : 0.001);   // tools/exp_sort_margin.py: 0.05 .. 1 mm 74.1 us, 2 mm 75.6, 8 mm 77.1
    // Root body turning about an axis parallel to gravity (z): its contribution to the joint-space dynamics is the
    // constant inertia about that axis (see ModelDev::root_spin). Axis of joint 0 in the world = first column of the
    // normalised tree_R[0]; the tilt tolerance (1e-9 rad) is far below the fp32 kernel's own rounding
    // (the URDF's "3.14159265359" already tilts the yaw axis by 2e-13 rad).
    {
        double a[3] = {(double)d.tree_R[0][0], (double)d.tree_R[0][3], (double)d.tree_R[0][6]};
        const bool vertical = fabs(a[0]) < 1e-9 && fabs(a[1]) < 1e-9 && fabs(fabs(a[2]) - 1.0) < 1e-9;
        bool carries_contact = false;
        for (int c = 0; c < m.n_contacts; ++c) carries_contact = carries_contact || m.contact_body[c] == 0;
        d.root_spin = (vertical && !tune.disable_root_fold) ? 1 : 0;
        (void)carries_contact;   // the proxies' positions are still computed from body 0's frame
        // |a x com|^2 with a = e_x in the normalised body frame: com_y^2 + com_z^2 ; a.I.a = Ixx
        const double cy = (double)d.com[0][1], cz = (double)d.com[0][2];
        d.root_mass_term = (T)(m.mass[0] * (cy * cy + cz * cz));
        d.root_inertia_term = (T)(double)d.inertia[0][0];
    }
    d.any_damping = 0;
    for (int i = 0; i < m.n_dof; ++i) if (m.damping[i] != 0.0) d.any_damping = 1;
}

// Structure signature of a model (os2r_device.cuh): which joints' tree_R is the identity / a turn about the joint axis /
// a turn about z in the normalised frames, and which components of tree_p / contact_pos are zero. |x| <= 1e-15 counts as
// zero (the URDFs' literal rpy "1.57" leaves products of the order 1e-19 in entries that are structurally 0: they are
// below half an ulp of the fp64 sum they would enter).
void model_signature(const os2r_model &m, const os2r_tuning &tune, uint32_t *sj, uint32_t *sc) {
    ModelDev<double> d;
    build_model_dev<double>(m, tune, d);
    auto z = [](double x) { return fabs(x) <= 1e-15; };
    auto one = [](double x) { return fabs(x - 1.0) <= 1e-15; };
    uint32_t J = 0, C = 0;
    for (int i = m.n_dof - 1; i >= 0; --i) {
        const double *R = d.tree_R[i], *p = d.tree_p[i];
        uint32_t kind = OS2R_TREE_GENERAL;
        if (i > 0) {   // joint 0 hangs off the world: its constant frame is used as it is
            if (one(R[0]) && z(R[1]) && z(R[2]) && z(R[3]) && one(R[4]) && z(R[5]) && z(R[6]) && z(R[7]) && one(R[8])) kind = OS2R_TREE_IDENTITY;
            else if (one(R[0]) && z(R[1]) && z(R[2]) && z(R[3]) && z(R[6]) && z(R[4] - R[8]) && z(R[5] + R[7])) kind = OS2R_TREE_XTURN;
            else if (one(R[8]) && z(R[2]) && z(R[5]) && z(R[6]) && z(R[7])) kind = OS2R_TREE_ZTURN;
            else if (one(R[4]) && z(R[1]) && z(R[3]) && z(R[5]) && z(R[7])) kind = OS2R_TREE_YTURN;
        }
        uint32_t pm = 0;
        for (int c = 0; c < 3; ++c) if (i == 0 || !z(p[c])) pm |= 1u << c;
        J = (J << 6) | (pm << 3) | kind;
    }
    for (int c = m.n_contacts - 1; c >= 0; --c) {
        uint32_t cm = 0;
        for (int k = 0; k < 3; ++k) if (!z(d.contact_pos[c][k])) cm |= 1u << k;
        C = (C << 6) | ((uint32_t)m.contact_body[c] << 3) | cm;
    }
    J |= (d.root_spin ? 2u : 1u) << 30;
    *sj = J; *sc = C;
}

}  // namespace

struct os2r_env {
    os2r_model model;
    os2r_task_cfg task;
    TaskDev taskdev;
    ModelDev<float> m32;
    ModelDev<double> m64;
    int precision = 32;
    int device = 0;
    int64_t n = 0;
    int64_t first_env_id = 0;
    uint64_t seed = 0;
    int rows = 0;
    uint32_t sig_j = 0, sig_c = 0;   // structure signature the step kernel is picked by
    // device memory
    void *real_block = nullptr;      // all T-typed SoA arrays, one allocation
    size_t real_count = 0;           // elements of T
    int32_t *steps = nullptr;
    uint32_t *episode = nullptr;
    int32_t *reset_id = nullptr;
    double *ret = nullptr;
    uint8_t *cls = nullptr;
    StatsDev *stats = nullptr;
    int sm_count = 0;
    int build = OS2R_BUILD_F32;      // which step kernel: fp32 (product) or fp64 (verification)
    int block = OS2R_BLOCK;          // threads per block of the step kernel for this batch size
    StateDev<float> s32;
    StateDev<double> s64;
    // host staging for os2r_step_host
    cudaStream_t host_stream = nullptr;
    float *pin_actions = nullptr, *pin_obs = nullptr, *pin_reward = nullptr, *pin_term = nullptr;
    uint8_t *pin_done = nullptr;
    int32_t *pin_info = nullptr;
    float *dev_actions = nullptr, *dev_obs = nullptr, *dev_reward = nullptr, *dev_term = nullptr;
    uint8_t *dev_done = nullptr;
    int32_t *dev_info = nullptr;
    bool host_io_ready = false;
    // single-block I/O of os2r_step_host_packed
    unsigned char *dev_block = nullptr, *pin_block = nullptr;
    size_t pin_block_bytes = 0;
    bool packed_pending = false, packed_block_pinned = false;   // a packed step enqueued by ..._begin, not yet completed
    void *packed_block = nullptr;
    int32_t packed_prefix = 0;
    int64_t launches = 0;
    uint64_t env_steps = 0;
};

namespace {

template <typename T>
void carve(os2r_env *h, StateDev<T> &S) {
    T *base = (T *)h->real_block;
    const int64_t N = h->n;
    const int n = h->model.n_dof, nc = h->model.n_contacts;
    size_t off = 0;
    auto take = [&](int count) { T *p = base + off; off += (size_t)count * N; return p; };
    S.q_hi = take(n); S.q_lo = take(n); S.qd = take(n); S.qd_lo = take(n);
    S.lam = take(h->rows); S.a_prev = take(2);
    S.mass_scale = take(n); S.damping = take(n); S.friction = take(n);
    S.mu = take(nc); S.gravity_z = take(1);
    S.steps = h->steps; S.episode = h->episode; S.reset_id = h->reset_id; S.ret = h->ret; S.cls = h->cls;
    S.n_envs = N; S.first_env_id = h->first_env_id; S.seed = h->seed;
}

size_t real_elems(const os2r_model &m, int64_t N) {
    const int n = m.n_dof, nc = m.n_contacts;
    return (size_t)(4 * n + (n + 3 * nc) + 2 + 3 * n + nc + 1) * (size_t)N;
}

int ensure_host_io(os2r_env *h) {
    if (h->host_io_ready) return 0;
    const int64_t N = h->n;
    const int D = h->task.obs_dim;
    // a BLOCKING stream: implicitly ordered after work on the legacy default stream (what torch uses by default),
    // so a reset / device-buffer step enqueued there cannot race with a following host-buffer step
    CK(cudaStreamCreate(&h->host_stream));
    CK(cudaMallocHost(&h->pin_actions, N * 2 * sizeof(float)));
    CK(cudaMallocHost(&h->pin_obs, N * D * sizeof(float)));
    CK(cudaMallocHost(&h->pin_term, N * D * sizeof(float)));
    CK(cudaMallocHost(&h->pin_reward, N * sizeof(float)));
    CK(cudaMallocHost(&h->pin_done, N));
    CK(cudaMallocHost(&h->pin_info, N * 2 * sizeof(int32_t)));
    CK(cudaMalloc(&h->dev_actions, N * 2 * sizeof(float)));
    CK(cudaMalloc(&h->dev_obs, N * D * sizeof(float)));
    CK(cudaMalloc(&h->dev_term, N * D * sizeof(float)));
    CK(cudaMalloc(&h->dev_reward, N * sizeof(float)));
    CK(cudaMalloc(&h->dev_done, N));
    CK(cudaMalloc(&h->dev_info, N * 2 * sizeof(int32_t)));
    h->host_io_ready = true;
    return 0;
}

template <typename T>
int get_real(os2r_env *h, const T *dev, int count, std::vector<double> &out) {
    std::vector<T> tmp((size_t)count * h->n);
    CK(cudaMemcpy(tmp.data(), dev, tmp.size() * sizeof(T), cudaMemcpyDeviceToHost));
    out.resize(tmp.size());
    for (size_t i = 0; i < tmp.size(); ++i) out[i] = (double)tmp[i];
    return 0;
}
template <typename T>
int put_real(os2r_env *h, T *dev, int count, const std::vector<double> &in) {
    std::vector<T> tmp((size_t)count * h->n);
    for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = (T)in[i];
    CK(cudaMemcpy(dev, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

template <typename T>
int get_state_impl(os2r_env *h, StateDev<T> &S, double *out) {
    const int n = h->model.n_dof, W = os2r_state_width(&h->model);
    const int64_t N = h->n;
    std::vector<double> hi, lo, qd, qdl, lam, ap;
    if (get_real(h, S.q_hi, n, hi) || get_real(h, S.q_lo, n, lo) || get_real(h, S.qd, n, qd) || get_real(h, S.qd_lo, n, qdl) ||
        get_real(h, S.lam, h->rows, lam) || get_real(h, S.a_prev, 2, ap)) return 1;
    for (int64_t e = 0; e < N; ++e) {
        double *row = out + e * W;
        for (int i = 0; i < n; ++i) { row[i] = hi[i * N + e] + lo[i * N + e]; row[n + i] = qd[i * N + e] + qdl[i * N + e]; }
        for (int r = 0; r < h->rows; ++r) row[2 * n + r] = lam[r * N + e];
        row[2 * n + h->rows] = ap[e];
        row[2 * n + h->rows + 1] = ap[N + e];
    }
    return 0;
}
template <typename T>
int set_state_impl(os2r_env *h, StateDev<T> &S, const double *in) {
    const int n = h->model.n_dof, W = os2r_state_width(&h->model);
    const int64_t N = h->n;
    std::vector<double> hi((size_t)n * N), lo((size_t)n * N), qd((size_t)n * N), qdl((size_t)n * N), lam((size_t)h->rows * N), ap((size_t)2 * N);
    for (int64_t e = 0; e < N; ++e) {
        const double *row = in + e * W;
        for (int i = 0; i < n; ++i) {
            const T h1 = (T)row[i];
            hi[i * N + e] = (double)h1;
            lo[i * N + e] = sizeof(T) == 4 ? (double)(T)(row[i] - (double)h1) : 0.0;
            const T v1 = (T)row[n + i];
            qd[i * N + e] = (double)v1;
            qdl[i * N + e] = sizeof(T) == 4 ? (double)(T)(row[n + i] - (double)v1) : 0.0;
        }
        for (int r = 0; r < h->rows; ++r) lam[r * N + e] = row[2 * n + r];
        ap[e] = row[2 * n + h->rows];
        ap[N + e] = row[2 * n + h->rows + 1];
    }
    CK(cudaMemset(h->cls, 0xFF, (size_t)N));   // sorting hint unknown for a state written from outside
    return put_real(h, S.q_hi, n, hi) || put_real(h, S.q_lo, n, lo) || put_real(h, S.qd, n, qd) || put_real(h, S.qd_lo, n, qdl) ||
           put_real(h, S.lam, h->rows, lam) || put_real(h, S.a_prev, 2, ap);
}
template <typename T>
int get_params_impl(os2r_env *h, StateDev<T> &S, double *out) {
    const int n = h->model.n_dof, nc = h->model.n_contacts, PW = os2r_params_width(&h->model);
    const int64_t N = h->n;
    std::vector<double> ms, dm, fr, mu, gz;
    if (get_real(h, S.mass_scale, n, ms) || get_real(h, S.damping, n, dm) || get_real(h, S.friction, n, fr) ||
        get_real(h, S.mu, nc, mu) || get_real(h, S.gravity_z, 1, gz)) return 1;
    for (int64_t e = 0; e < N; ++e) {
        double *row = out + e * PW;
        for (int i = 0; i < n; ++i) { row[i] = ms[i * N + e]; row[n + i] = dm[i * N + e]; row[2 * n + i] = fr[i * N + e]; }
        for (int c = 0; c < nc; ++c) row[3 * n + c] = mu[c * N + e];
        row[3 * n + nc] = gz[e];
    }
    return 0;
}
template <typename T>
int set_params_impl(os2r_env *h, StateDev<T> &S, const double *in) {
    const int n = h->model.n_dof, nc = h->model.n_contacts, PW = os2r_params_width(&h->model);
    const int64_t N = h->n;
    std::vector<double> ms((size_t)n * N), dm((size_t)n * N), fr((size_t)n * N), mu((size_t)nc * N), gz((size_t)N);
    for (int64_t e = 0; e < N; ++e) {
        const double *row = in + e * PW;
        for (int i = 0; i < n; ++i) { ms[i * N + e] = row[i]; dm[i * N + e] = row[n + i]; fr[i * N + e] = row[2 * n + i]; }
        for (int c = 0; c < nc; ++c) mu[c * N + e] = row[3 * n + c];
        gz[e] = row[3 * n + nc];
    }
    // a damping written from outside turns on the implicit-damping factorisation even for a model whose nominal
    // damping is zero (the step kernel is compiled with and without it, os2r_kernels.cu: DAMPED)
    for (size_t i = 0; i < dm.size(); ++i)
        if (dm[i] != 0.0) { h->m32.any_damping = 1; h->m64.any_damping = 1; break; }
    return put_real(h, S.mass_scale, n, ms) || put_real(h, S.damping, n, dm) || put_real(h, S.friction, n, fr) ||
           put_real(h, S.mu, nc, mu) || put_real(h, S.gravity_z, 1, gz);
}

int do_step(os2r_env *h, const StepIO &io, cudaStream_t stream) {
    cudaError_t e;
    if (h->precision == 32)
        e = launch_step<float>(h->build, h->model.n_dof, h->model.n_contacts, h->block, h->sig_j, h->sig_c, h->m32, h->taskdev, h->s32, io, h->stats, stream);
    else
        e = launch_step<double>(h->build, h->model.n_dof, h->model.n_contacts, h->block, h->sig_j, h->sig_c, h->m64, h->taskdev, h->s64, io, h->stats, stream);
    if (e != cudaSuccess) return fail("step kernel launch failed: %s", cudaGetErrorString(e));
    h->launches += 1;
    h->env_steps += (uint64_t)h->n;
    return 0;
}

}  // namespace

extern "C" {

int32_t os2r_abi_version(void) { return OS2R_ABI_VERSION; }
const char *os2r_last_error(void) { return g_err; }

int32_t os2r_state_width(const os2r_model *m) { return 2 * m->n_dof + (m->n_dof + 3 * m->n_contacts) + 2; }
int32_t os2r_params_width(const os2r_model *m) { return 3 * m->n_dof + m->n_contacts + 1; }

int32_t os2r_create(const os2r_model *model, const os2r_task_cfg *task, int64_t n_envs, int64_t first_env_id,
                    int32_t device, uint64_t seed, int32_t precision, os2r_env **out) {
    return os2r_create_tuned(model, task, n_envs, first_env_id, device, seed, precision, nullptr, out);
}

int32_t os2r_create_tuned(const os2r_model *model, const os2r_task_cfg *task, int64_t n_envs, int64_t first_env_id,
                          int32_t device, uint64_t seed, int32_t precision, const os2r_tuning *tuning, os2r_env **out) {
    if (!model || !task || !out) return fail("os2r_create: null argument");
    os2r_tuning tune;
    memset(&tune, 0, sizeof(tune));
    if (tuning) tune = *tuning;
    *out = nullptr;
    if (n_envs <= 0) return fail("os2r_create: n_envs must be positive (got %lld)", (long long)n_envs);
    if (model->n_dof < 2 || model->n_dof > OS2R_MAX_DOF) return fail("os2r_create: n_dof %d unsupported (2..5)", model->n_dof);
    if (!supported_shape(model->n_dof, model->n_contacts))
        return fail("os2r_create: no kernel is built for %d moving joints with %d contact proxies (built: 2..5 joints with 3 proxies, 5 joints with 4)", model->n_dof, model->n_contacts);
    if (precision != 32 && precision != 64) return fail("os2r_create: precision must be 32 or 64");
    if (!(model->pgs_tol >= 0.0)) return fail("os2r_create: pgs_tol must be >= 0");
    if (model->pgs_joint_sweeps < 0) return fail("os2r_create: pgs_joint_sweeps must be >= 0");
    if (task->obs_dim <= 0 || task->obs_dim > OS2R_MAX_OBS) return fail("os2r_create: obs_dim %d out of range", task->obs_dim);
    if (task->n_resets <= 0 || task->n_resets > OS2R_MAX_RESETS) return fail("os2r_create: n_resets %d out of range", task->n_resets);
    if (model->role_dof[OS2R_ROLE_HIP] < 0 || model->role_dof[OS2R_ROLE_KNEE] < 0) return fail("os2r_create: model lacks hip/knee joints");
    for (int i = 0; i < model->n_dof; ++i) if (model->axis[i] < 0 || model->axis[i] > 2) return fail("os2r_create: bad joint axis");
    for (int c = 0; c < model->n_contacts; ++c)
        if (model->contact_body[c] < 0 || model->contact_body[c] >= model->n_dof) return fail("os2r_create: bad contact body");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail("os2r_create: no CUDA device available (%s); this backend has no CPU fallback",
                    ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
    if (device < 0 || device >= ndev) return fail("os2r_create: device %d out of range (0..%d)", device, ndev - 1);
    DeviceGuard guard(device);
    int sm_count = 0;
    if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sm_count <= 0)
        return fail("os2r_create: cannot query the SM count of device %d", device);

    os2r_env *h = new os2r_env();
    h->sm_count = sm_count;
    h->build = precision == 64 ? OS2R_BUILD_F64 : OS2R_BUILD_F32;
    h->block = step_block_threads(h->build, n_envs, sm_count);
    if (tune.force_block == OS2R_BLOCK || (precision == 32 && tune.force_block == OS2R_BLOCK_WIDE)) h->block = tune.force_block;
    else if (tune.force_block != 0) { delete h; return fail("os2r_create: tuning.force_block %d unsupported (%d, or %d for the fp32 builds)", tune.force_block, OS2R_BLOCK, OS2R_BLOCK_WIDE); }
    h->model = *model; h->task = *task; h->precision = precision; h->device = device;
    h->n = n_envs; h->first_env_id = first_env_id; h->seed = seed;
    h->rows = model->n_dof + 3 * model->n_contacts;
    // the kernels specialised on the shipped models' structure run when the model has exactly that structure
    model_signature(*model, tune, &h->sig_j, &h->sig_c);
    if (tune.disable_specialisation) { h->sig_j = generic_joint_signature(model->n_dof); h->sig_c = generic_contact_signature(model->n_contacts); }
    build_model_dev<float>(*model, tune, h->m32);
    build_model_dev<double>(*model, tune, h->m64);
    memset(&h->taskdev, 0, sizeof(h->taskdev));
    h->taskdev.cfg = *task;
    for (int i = 0; i < model->n_dof; ++i) { h->taskdev.nominal_damping[i] = model->damping[i]; h->taskdev.nominal_friction[i] = model->friction[i]; }
    for (int c = 0; c < model->n_contacts; ++c) h->taskdev.nominal_mu[c] = model->contact_mu[c];
    for (int r = 0; r < OS2R_N_ROLES; ++r) h->taskdev.role_dof[r] = model->role_dof[r];
    for (int k = 0; k < task->obs_dim && k < OS2R_MAX_OBS; ++k) h->taskdev.obs_scale[k] = 2.0 / (task->obs_high[k] - task->obs_low[k]);
    h->taskdev.n_dof = model->n_dof; h->taskdev.n_contacts = model->n_contacts;

    const size_t esz = precision == 32 ? sizeof(float) : sizeof(double);
    h->real_count = real_elems(*model, n_envs);
    auto cleanup = [&](const char *what, cudaError_t e) {
        fail("os2r_create: %s failed: %s", what, cudaGetErrorString(e));
        os2r_destroy(h);
        return 1;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&h->real_block, h->real_count * esz)) != cudaSuccess) return cleanup("cudaMalloc(state)", e);
    if ((e = cudaMalloc(&h->steps, n_envs * sizeof(int32_t))) != cudaSuccess) return cleanup("cudaMalloc(steps)", e);
    if ((e = cudaMalloc(&h->episode, n_envs * sizeof(uint32_t))) != cudaSuccess) return cleanup("cudaMalloc(episode)", e);
    if ((e = cudaMalloc(&h->reset_id, n_envs * sizeof(int32_t))) != cudaSuccess) return cleanup("cudaMalloc(reset_id)", e);
    if ((e = cudaMalloc(&h->ret, n_envs * sizeof(double))) != cudaSuccess) return cleanup("cudaMalloc(ret)", e);
    if ((e = cudaMalloc(&h->cls, n_envs)) != cudaSuccess) return cleanup("cudaMalloc(cls)", e);
    if ((e = cudaMalloc(&h->stats, sizeof(StatsDev))) != cudaSuccess) return cleanup("cudaMalloc(stats)", e);
    if ((e = cudaMemset(h->stats, 0, sizeof(StatsDev))) != cudaSuccess) return cleanup("cudaMemset(stats)", e);
    e = prepare_step(h->build, model->n_dof, model->n_contacts, h->block, h->sig_j, h->sig_c);
    if (e != cudaSuccess) return cleanup("cudaFuncSetAttribute(step kernel shared memory)", e);
    if (precision == 32) { carve<float>(h, h->s32); e = launch_init<float>(h->taskdev, h->s32, model->gravity_z, 0); }
    else { carve<double>(h, h->s64); e = launch_init<double>(h->taskdev, h->s64, model->gravity_z, 0); }
    if (e != cudaSuccess) return cleanup("init kernel", e);
    h->launches += 1;
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return cleanup("init kernel sync", e);
    *out = h;
    return 0;
}

int32_t os2r_destroy(os2r_env *h) {
    if (!h) return 0;
    DeviceGuard guard(h->device);
    if (h->host_stream) cudaStreamSynchronize(h->host_stream);   // a packed step may still be reading / writing the staging buffers
    cudaFree(h->real_block); cudaFree(h->steps); cudaFree(h->episode); cudaFree(h->reset_id); cudaFree(h->ret); cudaFree(h->cls);
    cudaFree(h->stats);
    if (h->host_io_ready || h->host_stream) {
        cudaFreeHost(h->pin_actions); cudaFreeHost(h->pin_obs); cudaFreeHost(h->pin_term); cudaFreeHost(h->pin_reward);
        cudaFreeHost(h->pin_done); cudaFreeHost(h->pin_info);
        cudaFree(h->dev_actions); cudaFree(h->dev_obs); cudaFree(h->dev_term); cudaFree(h->dev_reward);
        cudaFree(h->dev_done); cudaFree(h->dev_info);
        cudaFree(h->dev_block); cudaFreeHost(h->pin_block);
        if (h->host_stream) cudaStreamDestroy(h->host_stream);
    }
    delete h;
    return 0;
}

int32_t os2r_seed(os2r_env *h, uint64_t seed) {
    if (!h) return fail("os2r_seed: null handle");
    h->seed = seed; h->s32.seed = seed; h->s64.seed = seed;
    return 0;
}

int32_t os2r_reset(os2r_env *h, const uint8_t *mask_dev, float *obs_dev, void *stream) {
    if (!h) return fail("os2r_reset: null handle");
    DeviceGuard guard(h->device);
    cudaError_t e;
    if (h->precision == 32) e = launch_reset<float>(h->model.n_dof, h->model.n_contacts, h->taskdev, h->s32, mask_dev, obs_dev, (cudaStream_t)stream);
    else e = launch_reset<double>(h->model.n_dof, h->model.n_contacts, h->taskdev, h->s64, mask_dev, obs_dev, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail("reset kernel launch failed: %s", cudaGetErrorString(e));
    h->launches += 1;
    return 0;
}

int32_t os2r_step(os2r_env *h, const float *actions_dev, float *obs_dev, float *reward_dev, uint8_t *done_dev,
                  float *terminal_obs_dev, int32_t *info_dev, void *stream) {
    if (!h) return fail("os2r_step: null handle");
    if (!actions_dev || !obs_dev || !reward_dev || !done_dev) return fail("os2r_step: actions/obs/reward/done must be non-null");
    DeviceGuard guard(h->device);
    StepIO io{};
    io.actions = actions_dev; io.obs = obs_dev; io.reward = reward_dev; io.done = done_dev;
    io.term_obs = terminal_obs_dev; io.info = info_dev;
    return do_step(h, io, (cudaStream_t)stream);
}

// true when `p` is page-locked host memory the GPU can DMA to/from directly (cudaHostAlloc / cudaHostRegister)
static bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int32_t os2r_step_host(os2r_env *h, const float *actions, float *obs, float *reward, uint8_t *done,
                       float *terminal_obs, int32_t *info) {
    if (!h) return fail("os2r_step_host: null handle");
    if (!actions || !obs || !reward || !done) return fail("os2r_step_host: actions/obs/reward/done must be non-null");
    if (h->packed_pending) return fail("os2r_step_host: a packed step is in flight (call os2r_step_host_packed_end first)");
    DeviceGuard guard(h->device);
    if (ensure_host_io(h)) return 1;
    const int64_t N = h->n;
    const int D = h->task.obs_dim;
    cudaStream_t st = h->host_stream;
    // Pageable caller buffers are staged through the handle's pinned area (one extra host memcpy each way);
    // page-locked caller buffers are DMA targets themselves.
    const bool pa = is_pinned(actions), po = is_pinned(obs), pr = is_pinned(reward), pd = is_pinned(done);
    const bool pt = terminal_obs && is_pinned(terminal_obs), pi = info && is_pinned(info);
    if (!pa) memcpy(h->pin_actions, actions, N * 2 * sizeof(float));
    CK(cudaMemcpyAsync(h->dev_actions, pa ? actions : h->pin_actions, N * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    StepIO io{};
    io.actions = h->dev_actions; io.obs = h->dev_obs; io.reward = h->dev_reward; io.done = h->dev_done;
    io.term_obs = terminal_obs ? h->dev_term : nullptr; io.info = info ? h->dev_info : nullptr;
    if (do_step(h, io, st)) return 1;
    CK(cudaMemcpyAsync(po ? obs : h->pin_obs, h->dev_obs, N * D * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(pr ? reward : h->pin_reward, h->dev_reward, N * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(pd ? done : h->pin_done, h->dev_done, N, cudaMemcpyDeviceToHost, st));
    if (info) CK(cudaMemcpyAsync(pi ? info : h->pin_info, h->dev_info, N * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (!po) memcpy(obs, h->pin_obs, N * D * sizeof(float));
    if (!pr) memcpy(reward, h->pin_reward, N * sizeof(float));
    if (!pd) memcpy(done, h->pin_done, N);
    if (info && !pi) memcpy(info, h->pin_info, N * 2 * sizeof(int32_t));
    if (terminal_obs) {
        // The terminal observation differs from `obs` only for envs that finished an episode, so it is
        // fetched lazily: a second D2H only on steps where some env is done (rare: episodes are long).
        bool any_done = false;
        for (int64_t e = 0; e < N && !any_done; ++e) any_done = done[e] != 0;
        if (any_done) {
            CK(cudaMemcpyAsync(pt ? terminal_obs : h->pin_term, h->dev_term, N * D * sizeof(float), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (!pt) memcpy(terminal_obs, h->pin_term, N * D * sizeof(float));
        } else {
            memcpy(terminal_obs, obs, N * D * sizeof(float));
        }
    }
    return 0;
}

static void packed_layout(const os2r_env *h, int32_t prefix_records, os2r_packed_layout *L) {
    const int64_t N = h->n;
    const int D = h->task.obs_dim;
    auto align16 = [](int64_t x) { return (x + 15) & ~(int64_t)15; };
    L->obs = 0;
    L->reward = align16(L->obs + N * D * 4);
    L->done = align16(L->reward + N * 4);
    L->reset_id = align16(L->done + N);
    L->term_count = align16(L->reset_id + N);
    L->term_records = L->term_count + 16;
    L->record_words = D + 2;
    L->prefix_records = prefix_records;
    L->total_bytes = L->term_records + (int64_t)prefix_records * L->record_words * 4;
}

int32_t os2r_packed_layout_get(const os2r_env *h, int32_t prefix_records, os2r_packed_layout *out) {
    if (!h || !out) return fail("os2r_packed_layout_get: null argument");
    if (prefix_records < 0 || prefix_records > h->n) return fail("os2r_packed_layout_get: prefix_records must be in [0, n_envs]");
    packed_layout(h, prefix_records, out);
    return 0;
}

int32_t os2r_step_host_packed_begin(os2r_env *h, const float *actions, void *block, int32_t prefix_records) {
    if (!h || !actions || !block) return fail("os2r_step_host_packed_begin: null argument");
    if (prefix_records < 0 || prefix_records > h->n) return fail("os2r_step_host_packed_begin: prefix_records must be in [0, n_envs]");
    if (h->packed_pending) return fail("os2r_step_host_packed_begin: the previous packed step has not been completed (call os2r_step_host_packed_end)");
    DeviceGuard guard(h->device);
    if (ensure_host_io(h)) return 1;
    const int64_t N = h->n;
    os2r_packed_layout L, Lfull;
    packed_layout(h, prefix_records, &L);
    packed_layout(h, (int32_t)N, &Lfull);        // the device block can hold a record for every env
    if (!h->dev_block) CK(cudaMalloc(&h->dev_block, (size_t)Lfull.total_bytes));
    cudaStream_t st = h->host_stream;
    const bool pa = is_pinned(actions), pb = is_pinned(block);
    if (!pb && h->pin_block_bytes < (size_t)L.total_bytes) {
        if (h->pin_block) cudaFreeHost(h->pin_block);
        h->pin_block = nullptr; h->pin_block_bytes = 0;
        CK(cudaMallocHost(&h->pin_block, (size_t)L.total_bytes));
        h->pin_block_bytes = (size_t)L.total_bytes;
    }
    if (!pa) memcpy(h->pin_actions, actions, N * 2 * sizeof(float));
    CK(cudaMemcpyAsync(h->dev_actions, pa ? actions : h->pin_actions, N * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(h->dev_block + L.term_count, 0, 16, st));
    StepIO io{};
    io.actions = h->dev_actions;
    io.obs = (float *)(h->dev_block + L.obs);
    io.reward = (float *)(h->dev_block + L.reward);
    io.done = h->dev_block + L.done;
    io.reset_id8 = h->dev_block + L.reset_id;
    io.term_count = (int32_t *)(h->dev_block + L.term_count);
    io.term_records = (int32_t *)(h->dev_block + L.term_records);
    io.term_cap = (int32_t)N;
    if (do_step(h, io, st)) return 1;
    unsigned char *dst = pb ? (unsigned char *)block : h->pin_block;
    CK(cudaMemcpyAsync(dst, h->dev_block, (size_t)L.total_bytes, cudaMemcpyDeviceToHost, st));
    h->packed_pending = true;
    h->packed_block = block;
    h->packed_block_pinned = pb;
    h->packed_prefix = prefix_records;
    return 0;
}

int32_t os2r_step_host_packed_end(os2r_env *h, int32_t *n_terminal) {
    if (!h) return fail("os2r_step_host_packed_end: null handle");
    if (!h->packed_pending) return fail("os2r_step_host_packed_end: no packed step in flight");
    DeviceGuard guard(h->device);
    os2r_packed_layout L;
    packed_layout(h, h->packed_prefix, &L);
    h->packed_pending = false;
    CK(cudaStreamSynchronize(h->host_stream));
    if (!h->packed_block_pinned) memcpy(h->packed_block, h->pin_block, (size_t)L.total_bytes);
    if (n_terminal) *n_terminal = *(const int32_t *)((const unsigned char *)h->packed_block + L.term_count);
    return 0;
}

int32_t os2r_step_host_packed(os2r_env *h, const float *actions, void *block, int32_t prefix_records, int32_t *n_terminal) {
    if (os2r_step_host_packed_begin(h, actions, block, prefix_records)) return 1;
    return os2r_step_host_packed_end(h, n_terminal);
}

int32_t os2r_fetch_terminal_records(os2r_env *h, int32_t first, int32_t count, int32_t *records_host) {
    if (!h || !records_host) return fail("os2r_fetch_terminal_records: null argument");
    if (!h->dev_block) return fail("os2r_fetch_terminal_records: no packed step has run yet");
    if (first < 0 || count < 0 || (int64_t)first + count > h->n) return fail("os2r_fetch_terminal_records: range out of bounds");
    DeviceGuard guard(h->device);
    os2r_packed_layout L;
    packed_layout(h, 0, &L);
    const size_t rec = (size_t)L.record_words * 4;
    CK(cudaMemcpyAsync(records_host, h->dev_block + L.term_records + (size_t)first * rec, (size_t)count * rec,
                       cudaMemcpyDeviceToHost, h->host_stream));
    CK(cudaStreamSynchronize(h->host_stream));
    return 0;
}

int32_t os2r_get_state(os2r_env *h, double *state_host) {
    if (!h || !state_host) return fail("os2r_get_state: null argument");
    DeviceGuard guard(h->device);
    CK(cudaDeviceSynchronize());
    return h->precision == 32 ? get_state_impl<float>(h, h->s32, state_host) : get_state_impl<double>(h, h->s64, state_host);
}
int32_t os2r_set_state(os2r_env *h, const double *state_host) {
    if (!h || !state_host) return fail("os2r_set_state: null argument");
    DeviceGuard guard(h->device);
    CK(cudaDeviceSynchronize());
    return h->precision == 32 ? set_state_impl<float>(h, h->s32, state_host) : set_state_impl<double>(h, h->s64, state_host);
}
int32_t os2r_get_params(os2r_env *h, double *params_host) {
    if (!h || !params_host) return fail("os2r_get_params: null argument");
    DeviceGuard guard(h->device);
    CK(cudaDeviceSynchronize());
    return h->precision == 32 ? get_params_impl<float>(h, h->s32, params_host) : get_params_impl<double>(h, h->s64, params_host);
}
int32_t os2r_set_params(os2r_env *h, const double *params_host) {
    if (!h || !params_host) return fail("os2r_set_params: null argument");
    DeviceGuard guard(h->device);
    CK(cudaDeviceSynchronize());
    return h->precision == 32 ? set_params_impl<float>(h, h->s32, params_host) : set_params_impl<double>(h, h->s64, params_host);
}
int32_t os2r_get_episode(os2r_env *h, int32_t *steps_host, double *returns_host, int32_t *reset_ids_host,
                         uint32_t *episodes_host) {
    if (!h) return fail("os2r_get_episode: null handle");
    DeviceGuard guard(h->device);
    CK(cudaDeviceSynchronize());
    if (steps_host) CK(cudaMemcpy(steps_host, h->steps, h->n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (returns_host) CK(cudaMemcpy(returns_host, h->ret, h->n * sizeof(double), cudaMemcpyDeviceToHost));
    if (reset_ids_host) CK(cudaMemcpy(reset_ids_host, h->reset_id, h->n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (episodes_host) CK(cudaMemcpy(episodes_host, h->episode, h->n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return 0;
}
int32_t os2r_set_episode(os2r_env *h, const int32_t *steps_host, const double *returns_host,
                         const int32_t *reset_ids_host, const uint32_t *episodes_host) {
    if (!h) return fail("os2r_set_episode: null handle");
    DeviceGuard guard(h->device);
    CK(cudaDeviceSynchronize());
    if (reset_ids_host)
        for (int64_t e = 0; e < h->n; ++e)
            if (reset_ids_host[e] < 0 || reset_ids_host[e] >= h->task.n_resets)
                return fail("os2r_set_episode: reset id %d of env %lld out of range (0..%d)", reset_ids_host[e], (long long)e, h->task.n_resets - 1);
    if (steps_host) CK(cudaMemcpy(h->steps, steps_host, h->n * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (returns_host) CK(cudaMemcpy(h->ret, returns_host, h->n * sizeof(double), cudaMemcpyHostToDevice));
    if (reset_ids_host) CK(cudaMemcpy(h->reset_id, reset_ids_host, h->n * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (episodes_host) CK(cudaMemcpy(h->episode, episodes_host, h->n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return 0;
}

int32_t os2r_set_randomization(os2r_env *h, const os2r_task_cfg *cfg) {
    if (!h || !cfg) return fail("os2r_set_randomization: null argument");
    if (!(cfg->mass_lo > 0.0) || cfg->mass_hi < cfg->mass_lo) return fail("os2r_set_randomization: mass range must satisfy 0 < lo <= hi");
    if (cfg->fric_lo < 0.0 || cfg->fric_hi < cfg->fric_lo) return fail("os2r_set_randomization: friction range must satisfy 0 <= lo <= hi");
    if (cfg->damp_lo < 0.0 || cfg->damp_hi < cfg->damp_lo) return fail("os2r_set_randomization: damping range must satisfy 0 <= lo <= hi");
    if (cfg->mu_lo < 0.0 || cfg->mu_hi < cfg->mu_lo || cfg->mu_link < 0.0) return fail("os2r_set_randomization: mu range must satisfy 0 <= lo <= hi");
    if (cfg->grav_std < 0.0 || cfg->gravity_redraw_resets < 0) return fail("os2r_set_randomization: grav_std / gravity_redraw_resets must be >= 0");
    DeviceGuard guard(h->device);
    CK(cudaDeviceSynchronize());   // the task configuration is a kernel parameter: steps already enqueued keep the old one
    os2r_task_cfg &t = h->taskdev.cfg;
    t.mass_lo = cfg->mass_lo; t.mass_hi = cfg->mass_hi; t.fric_lo = cfg->fric_lo; t.fric_hi = cfg->fric_hi;
    t.damp_lo = cfg->damp_lo; t.damp_hi = cfg->damp_hi; t.mu_lo = cfg->mu_lo; t.mu_hi = cfg->mu_hi; t.mu_link = cfg->mu_link;
    t.grav_mean = cfg->grav_mean; t.grav_std = cfg->grav_std;
    t.reset_randomized = cfg->reset_randomized; t.randomize_params = cfg->randomize_params;
    t.randomize_gravity = cfg->randomize_gravity; t.gravity_redraw_resets = cfg->gravity_redraw_resets;
    t.simple_sample_reset = cfg->simple_sample_reset;   // derived from reset_randomized (`simple` mode + NoRandomizer)
    h->task = t;
    // a damping range on a model with damped joints keeps the implicit-damping build; nothing else depends on the ranges
    return 0;
}

int32_t os2r_stats_read(os2r_env *h, os2r_stats *out, int32_t clear) {
    if (!h || !out) return fail("os2r_stats_read: null argument");
    DeviceGuard guard(h->device);
    CK(cudaDeviceSynchronize());
    StatsDev s;
    CK(cudaMemcpy(&s, h->stats, sizeof(s), cudaMemcpyDeviceToHost));
    out->env_steps = h->env_steps;
    out->episodes = s.episodes; out->done_task = s.done_task; out->done_timelimit = s.done_timelimit;
    out->nonfinite_resets = s.nonfinite_resets; out->sum_return = s.sum_return; out->sum_length = s.sum_length;
    if (clear) { CK(cudaMemset(h->stats, 0, sizeof(StatsDev))); h->env_steps = 0; }
    return 0;
}

int32_t os2r_stats_write(os2r_env *h, const os2r_stats *in) {
    if (!h || !in) return fail("os2r_stats_write: null argument");
    DeviceGuard guard(h->device);
    CK(cudaDeviceSynchronize());
    StatsDev s;
    s.episodes = in->episodes; s.done_task = in->done_task; s.done_timelimit = in->done_timelimit;
    s.nonfinite_resets = in->nonfinite_resets; s.sum_return = in->sum_return; s.sum_length = in->sum_length;
    CK(cudaMemcpy(h->stats, &s, sizeof(s), cudaMemcpyHostToDevice));
    h->env_steps = in->env_steps;
    return 0;
}

int32_t os2r_host_action_buffer(os2r_env *h, float **actions_out) {
    if (!h || !actions_out) return fail("os2r_host_action_buffer: null argument");
    DeviceGuard guard(h->device);
    if (ensure_host_io(h)) return 1;
    *actions_out = h->pin_actions;
    return 0;
}

int64_t os2r_num_envs(const os2r_env *h) { return h ? h->n : 0; }
int32_t os2r_obs_dim(const os2r_env *h) { return h ? h->task.obs_dim : 0; }
int64_t os2r_kernel_launches(const os2r_env *h) { return h ? h->launches : 0; }

int32_t os2r_model_signature(const os2r_model *model, uint32_t *joints, uint32_t *contacts, int32_t *specialised) {
    if (!model) return fail("os2r_model_signature: null model");
    if (model->n_dof < 1 || model->n_dof > OS2R_MAX_DOF || model->n_contacts < 0 || model->n_contacts > OS2R_MAX_CONTACTS)
        return fail("os2r_model_signature: model shape out of range");
    uint32_t sj = 0, sc = 0, kj = 0, kc = 0;
    os2r_tuning tune;
    memset(&tune, 0, sizeof(tune));
    model_signature(*model, tune, &sj, &sc);
    if (joints) *joints = sj;
    if (contacts) *contacts = sc;
    if (specialised) *specialised = (shipped_signature(model->n_dof, model->n_contacts, &kj, &kc) && kj == sj && kc == sc) ? 1 : 0;
    return 0;
}

int32_t os2r_kernel_info(const os2r_env *h, int32_t *block_threads, int32_t *grid_blocks, int32_t *regs_per_thread,
                         int32_t *local_bytes_per_thread, int32_t *resident_blocks_per_sm, int32_t *envs_per_thread) {
    if (!h) return fail("os2r_kernel_info: null handle");
    DeviceGuard guard(h->device);
    cudaFuncAttributes a;
    int resident = 0, epb = h->block;
    cudaError_t e = step_kernel_attributes(h->build, h->model.n_dof, h->model.n_contacts, h->block, h->m32.any_damping != 0, h->sig_j, h->sig_c, &a, &resident, &epb);
    if (e != cudaSuccess) return fail("cudaFuncGetAttributes failed: %s", cudaGetErrorString(e));
    if (block_threads) *block_threads = h->block;
    if (grid_blocks) *grid_blocks = (int32_t)((h->n + epb - 1) / epb);
    if (envs_per_thread) *envs_per_thread = epb / h->block;
    if (regs_per_thread) *regs_per_thread = a.numRegs;
    if (local_bytes_per_thread) *local_bytes_per_thread = (int32_t)a.localSizeBytes;
    if (resident_blocks_per_sm) *resident_blocks_per_sm = resident;
    return 0;
}

int32_t os2r_debug_counters(int32_t device, uint64_t *out8, int32_t clear) {
    if (!out8) return fail("os2r_debug_counters: null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail("os2r_debug_counters: no such CUDA device %d", device);
    DeviceGuard guard(device);
    CK(cudaDeviceSynchronize());
    unsigned long long tmp[8];
    cudaError_t e = read_check_counters(tmp, clear != 0);
    if (e != cudaSuccess) return fail("os2r_debug_counters: %s", cudaGetErrorString(e));
    for (int i = 0; i < 8; ++i) out8[i] = tmp[i];
    return 0;
}

int32_t os2r_measure_fp32_peak(int32_t device, double *tflops_out, double *sm_clock_mhz_out) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail("os2r_measure_fp32_peak: no such CUDA device %d", device);
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    float *out = nullptr;
    CK(cudaMalloc(&out, (size_t)blocks * 256 * sizeof(float)));
    cudaEvent_t t0, t1;
    CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(t0, 0));
        cudaError_t e = launch_fma_peak(out, blocks, iters, 0);
        if (e != cudaSuccess) return fail("fma peak launch failed: %s", cudaGetErrorString(e));
        CK(cudaEventRecord(t1, 0));
        CK(cudaEventSynchronize(t1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, t0, t1));
        const double flops = 2.0 * 64.0 * (double)iters * 256.0 * blocks;   // 8 accumulators x 8 unroll FMAs
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1); cudaFree(out);
    if (tflops_out) *tflops_out = best;
    if (sm_clock_mhz_out) *sm_clock_mhz_out = prop.clockRate / 1000.0;
    return 0;
}

}  // extern "C"
