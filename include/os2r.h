/*
 * os2r.h — C-ABI of the B200-native monopod step path (libos2r.so).
 *
 * This is the drop-in boundary for the ONE hot path of gym-os2r: `env.step()` for N envs
 * (reference: gym_os2r/runtimes/gazebo_runtime.py:65-97 = 10x {task.set_action; gazebo.run()}
 * + task.get_observation / get_reward / is_done, and the SubprocVecEnv auto-reset,
 * gym_os2r/common/vec_env/subproc_vec_env.py:14-21).  The reference has no FFI of its own
 * (it is pure Python over the un-vendored gym-ignition/ScenarIO/DART stack), so each entry
 * point below cites the reference *Python* interface it replaces.
 *
 * Conventions
 *  - plain C types only; no torch / CUDA types in signatures (streams travel as void*).
 *  - every function returns 0 on success, non-zero on failure; os2r_last_error() gives the
 *    thread-local message (mirrors the reference's bool-return -> assert/RuntimeError idiom,
 *    e.g. gym_os2r/tasks/monopod.py:225-229, gym_os2r/randomizers/monopod.py:60-61).
 *  - the library owns structure-of-arrays env state in HBM; the caller owns I/O buffers.
 *  - one handle per GPU; a handle is not thread-safe (the reference is single-threaded per env).
 *  - there is NO CPU fallback: creating a handle without a CUDA device fails loudly.
 */
#ifndef OS2R_H
#define OS2R_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OS2R_ABI_VERSION 9

#define OS2R_MAX_DOF 5
#define OS2R_MAX_CONTACTS 6
#define OS2R_MAX_OBS 12
#define OS2R_MAX_RESETS 8
#define OS2R_MAX_ROWS (OS2R_MAX_DOF + 3 * OS2R_MAX_CONTACTS)

/* joint roles (index into os2r_model.role_dof) — names follow the URDF joints
 * (gym_os2r/models/models/monopod/monopod.urdf:50,95,138,200,261) */
enum {
    OS2R_ROLE_HIP = 0,
    OS2R_ROLE_KNEE = 1,
    OS2R_ROLE_PITCH = 2,
    OS2R_ROLE_YAW = 3,
    OS2R_ROLE_BOOM_CONNECTOR = 4,
    OS2R_N_ROLES = 5
};

/* observation column kinds (gym_os2r/tasks/monopod.py:238-272) */
enum {
    OS2R_OBS_POS = 0,          /* joint position, affine-normalised to [-1,1] by its limits       */
    OS2R_OBS_POS_PERIODIC = 1, /* joint position wrapped to [-pi,pi) then normalised by +-pi      */
    OS2R_OBS_VEL = 2,          /* joint velocity, tanh(0.05 v)                                    */
    OS2R_OBS_TORQUE = 3        /* previous normalised action (observing_measured_torque)          */
};

/* built-in reward ids (gym_os2r/rewards/__init__.py:66-207) */
enum {
    OS2R_REWARD_CUSTOM = 0,      /* reward column left 0; host computes it (user RewardBase subclass) */
    OS2R_REWARD_BALANCING_V1 = 1,/* == StandingV1                                                      */
    OS2R_REWARD_BALANCING_V2 = 2,
    OS2R_REWARD_BALANCING_V3 = 3,
    OS2R_REWARD_HOPPING_V1 = 4,
    OS2R_REWARD_STRAIGHT_V1 = 5
};

/* Constant kinematic-tree / inertia tables, produced once on the host from the URDF
 * (replaces gym_os2r/models/monopod.py:10-38 `world.insert_model(urdf)`).
 * Bodies are the MOVING links only, root -> tip, fixed joints lumped into their parent. */
typedef struct os2r_model {
    int32_t n_dof;
    int32_t n_contacts;
    int32_t substeps;                 /* physics iterations per env step (10)                    */
    int32_t pgs_iters;                /* projected Gauss-Seidel sweeps per physics iteration     */
    int32_t axis[OS2R_MAX_DOF];       /* joint axis in the child frame: 0=x 1=y 2=z              */
    int32_t role_dof[OS2R_N_ROLES];   /* chain index of hip/knee/pitch/yaw/boom_connector or -1  */
    int32_t contact_body[OS2R_MAX_CONTACTS];
    int32_t pgs_joint_sweeps;         /* the joint-friction rows are relaxed in the first pgs_joint_sweeps sweeps of a
                                         physics iteration only, the contact rows in every sweep (default 1: their
                                         impulses are bounded by friction*dt ~ 1e-6 N m s and warm-started, the sweeps
                                         after the first iterate on the contacts); 0: in every sweep              */
    double tree_R[OS2R_MAX_DOF][9];   /* row-major; parent coords = R * child coords at q = 0    */
    double tree_p[OS2R_MAX_DOF][3];   /* joint origin in the parent (moving body or world) frame */
    double mass[OS2R_MAX_DOF];
    double com[OS2R_MAX_DOF][3];      /* body frame                                              */
    double inertia[OS2R_MAX_DOF][6];  /* about the COM, body axes: xx yy zz xy xz yz             */
    double damping[OS2R_MAX_DOF];     /* N m s / rad  (URDF <dynamics damping>)                  */
    double friction[OS2R_MAX_DOF];    /* N m          (URDF <dynamics friction>, Coulomb)        */
    double contact_pos[OS2R_MAX_CONTACTS][3]; /* sphere centre in its body frame                */
    double contact_radius[OS2R_MAX_CONTACTS];
    double contact_mu[OS2R_MAX_CONTACTS];     /* nominal effective mu = min(link, ground)       */
    double gravity_z;                 /* -9.8                                                    */
    double dt;                        /* physics step, 1e-4 s                                    */
    double erp, max_erv;              /* contact error reduction: min(depth*erp/dt, max_erv)     */
    double cfm_contact, cfm_joint;    /* relative constraint force mixing on the diagonal        */
    double max_torque[2];             /* hip, knee (settings.yaml:13-14)                         */
    double pgs_tol;                   /* > 0: an env's sweeps end after the first sweep whose velocity
                                         change sqrt(dv^T M dv) is <= pgs_tol (kinetic-energy norm,
                                         sqrt(kg) m/s); pgs_iters is then the cap. 0: always pgs_iters  */
} os2r_model;

/* Task / reward / termination / reset configuration (replaces MonopodTask.create_spaces,
 * gym_os2r/tasks/monopod.py:105-200, and the randomizers' reset logic,
 * gym_os2r/randomizers/monopod.py:67-135, monopod_no_rand.py:26-98). */
typedef struct os2r_task_cfg {
    int32_t obs_dim;
    int32_t normalized;               /* 1: tasks/monopod.py, 0: tasks/monopod_no_norm.py        */
    int32_t reward_id;
    int32_t max_episode_steps;        /* gym TimeLimit; 0 disables                               */
    int32_t auto_reset;               /* SubprocVecEnv semantics (subproc_vec_env.py:17-20)      */
    int32_t n_resets;
    int32_t reset_randomized;         /* 1: MonopodEnvRandomizer pose noise; 0: NoRandomizer     */
    int32_t randomize_params;         /* 1: per-reset mass/friction/damping/mu draws             */
    int32_t randomize_gravity;        /* 1: g_z ~ N(mean,std) drawn once per env at creation     */
    int32_t simple_sample_reset;      /* `simple` mode + NoRandomizer: hip,knee ~ obs space      */
    int32_t gravity_redraw_resets;    /* K > 0: g_z is re-drawn at every K-th reset of an env
                                         (MonopodEnvRandomizer(num_physics_rollouts=K),
                                         randomizers/monopod.py:36,56-61,371); 0: drawn once       */
    int32_t _pad1;
    int32_t reward_pitch_col, reward_yawvel_col, reward_hip_col, reward_knee_col; /* or -1       */
    int32_t obs_kind[OS2R_MAX_OBS];
    int32_t obs_index[OS2R_MAX_OBS];  /* chain dof index (POS/VEL) or action index (TORQUE)      */
    int32_t reset_laying[OS2R_MAX_RESETS];
    double obs_low[OS2R_MAX_OBS], obs_high[OS2R_MAX_OBS];   /* normalisation limits              */
    double done_low[OS2R_MAX_OBS], done_high[OS2R_MAX_OBS]; /* raw-unit termination thresholds:
                                         done iff raw < done_low or raw > done_high              */
    double reset_pitch[OS2R_MAX_RESETS];
    double simple_lo[2], simple_hi[2];/* hip,knee sample range for simple_sample_reset
                                         (monopod_no_rand.py:84: observation_space.sample())     */
    double ik_upper_leg, ik_lower_leg, ik_pivot_height, ik_boom, ik_hip_offset, ik_clip; /* mm  */
    double mass_lo, mass_hi;          /* coefficient ~ U(lo,hi)   (randomizers/monopod.py:183)   */
    double fric_lo, fric_hi;          /* absolute    ~ U(lo,hi)   (:191)                         */
    double damp_lo, damp_hi;          /* coefficient ~ U(lo,hi), zeros skipped (:199)            */
    double mu_lo, mu_hi, mu_link;     /* mu = mu_link * U(lo,hi)  (:207)                         */
    double grav_mean, grav_std;       /* (:58)                                                   */
} os2r_task_cfg;

/* Episode statistics accumulated on the device (the only quantity that is ever reduced
 * across GPUs; no reference counterpart beyond the prints in examples/fixed_hip.py:57). */
typedef struct os2r_stats {
    uint64_t env_steps;        /* env steps executed since creation / last clear  */
    uint64_t episodes;         /* episodes finished                               */
    uint64_t done_task;        /* ... because obs left the reset space            */
    uint64_t done_timelimit;   /* ... because of max_episode_steps                */
    uint64_t nonfinite_resets; /* envs force-reset because state became NaN/Inf   */
    double sum_return;         /* sum of finished episodes' returns               */
    double sum_length;         /* sum of finished episodes' lengths               */
} os2r_stats;

typedef struct os2r_env os2r_env; /* opaque handle */

int32_t os2r_abi_version(void);
const char *os2r_last_error(void);

/* Doubles per env in the packed state used by os2r_get_state / os2r_set_state:
 * [q(n_dof), qd(n_dof), lambda(n_dof + 3*n_contacts), a_prev(2)]; q/qd in chain order,
 * lambda = warm-start constraint impulses (joint-friction rows, then per contact n,t1,t2),
 * a_prev = the last applied normalised action (action_history[0], tasks/monopod.py:232-235). */
int32_t os2r_state_width(const os2r_model *model);
/* Doubles per env in os2r_get_params / os2r_set_params:
 * [mass_scale(n_dof), damping(n_dof), friction(n_dof), mu(n_contacts), gravity_z]. */
int32_t os2r_params_width(const os2r_model *model);

/* Create n_envs monopods on CUDA device `device`. `first_env_id` is the global index of this
 * shard's env 0 (per-env RNG streams are keyed by the global id, so results do not depend on
 * how envs are sharded over GPUs). precision: 32 = fp32 kernel with compensated positions
 * (the product path), 64 = fp64 kernel (verification path).
 * Replaces GazeboRuntime.__init__ (runtimes/gazebo_runtime.py:34-58). */
int32_t os2r_create(const os2r_model *model, const os2r_task_cfg *task, int64_t n_envs,
                    int64_t first_env_id, int32_t device, uint64_t seed, int32_t precision,
                    os2r_env **out);
/* Scheduling / debugging knobs that never change what an env computes (they replace the getenv switches of ABI v7).
 * Zero-initialise for the defaults. */
typedef struct os2r_tuning {
    double sort_margin;        /* ground clearance (m) below which a contact proxy counts as "near" for the lane sort;
                                  <= 0: default 0.001                                                              */
    int32_t force_block;       /* threads per block of the step kernel (64, or 224 for the fp32 builds); 0: chosen
                                  from the batch size                                                             */
    int32_t disable_specialisation; /* 1: run the all-general step kernel even when the model has the structure of a
                                  shipped URDF (verification of the specialised kernels)                          */
    int32_t disable_root_fold; /* 1: run the general per-body code for the yaw pivot (verification of the fold)  */
    int32_t _pad;
} os2r_tuning;
/* os2r_create with explicit tuning (NULL = defaults = os2r_create). */
int32_t os2r_create_tuned(const os2r_model *model, const os2r_task_cfg *task, int64_t n_envs,
                          int64_t first_env_id, int32_t device, uint64_t seed, int32_t precision,
                          const os2r_tuning *tuning, os2r_env **out);
int32_t os2r_destroy(os2r_env *env);

/* Replace the randomisation ranges / switches of a live handle (mass_lo..grav_std, reset_randomized,
 * randomize_params, randomize_gravity, simple_sample_reset, gravity_redraw_resets of `cfg`; everything else in `cfg`
 * is ignored). Takes
 * effect from the next reset of each env. Replaces editing MonopodRandomizersMixin's randomization_config
 * (randomizers/monopod.py:182-215). */
int32_t os2r_set_randomization(os2r_env *env, const os2r_task_cfg *cfg);

/* env.seed(seed): re-key the RNG streams (runtime seed(), tests/tests_general.py:20). */
int32_t os2r_seed(os2r_env *env, uint64_t seed);

/* Reset the envs whose mask byte is non-zero (all when mask_dev == NULL); draws the reset
 * pose (and physics parameters when randomize_params) and writes the reset observation into
 * obs_dev[N, obs_dim] for those envs (obs_dev may be NULL).
 * Replaces randomize_task + reset_task + get_observation
 * (randomizers/monopod.py:67-135, tasks/monopod.py:300-318,238-272). */
int32_t os2r_reset(os2r_env *env, const uint8_t *mask_dev, float *obs_dev, void *stream);

/* One env step for all N envs, device buffers:
 * actions_dev[N,2] in [-1,1] -> obs_dev[N,obs_dim], reward_dev[N], done_dev[N] (0/1),
 * optional terminal_obs_dev[N,obs_dim] (observation before auto-reset) and
 * info_dev[N,2] int32 = {reset_orientation id after the step, done cause bits
 * (1 task, 2 TimeLimit, 4 non-finite)}.
 * Replaces GazeboRuntime.step (runtimes/gazebo_runtime.py:65-97) and, with auto_reset, the
 * SubprocVecEnv worker (common/vec_env/subproc_vec_env.py:14-21). Asynchronous on `stream`. */
int32_t os2r_step(os2r_env *env, const float *actions_dev, float *obs_dev, float *reward_dev,
                  uint8_t *done_dev, float *terminal_obs_dev, int32_t *info_dev, void *stream);

/* Same step with HOST buffers (pageable or pinned): stages actions through pinned memory,
 * H2D, step, D2H, and returns when the outputs are valid. terminal_obs/info may be NULL.
 * This is what a numpy-facing caller (VecEnv.step, common/vec_env/vec_env.py:162-174) uses. */
int32_t os2r_step_host(os2r_env *env, const float *actions, float *obs, float *reward,
                       uint8_t *done, float *terminal_obs, int32_t *info);

/* The numpy-facing step at full batch size: everything a VecEnv.step returns comes back in ONE page-locked host
 * block through ONE device-to-host copy — obs, reward, done, one byte of reset-orientation id per env, and, instead
 * of a dense [N, obs_dim] terminal-observation array, one record per env that FINISHED an episode in this step
 * (info['terminal_observation'] exists only for those, subproc_vec_env.py:17-20). Byte offsets from the block start: */
typedef struct os2r_packed_layout {
    int64_t obs;            /* float32 [N, obs_dim]                                                    */
    int64_t reward;         /* float32 [N]                                                             */
    int64_t done;           /* uint8   [N] 0/1                                                         */
    int64_t reset_id;       /* uint8   [N] index into the task's reset_positions after the step        */
    int64_t term_count;     /* int32   [1] number of envs that finished an episode in this step        */
    int64_t term_records;   /* int32   [prefix_records][record_words]: {env index, done-cause bits,
                               terminal observation as float32 bit patterns [obs_dim]}, unordered      */
    int64_t total_bytes;    /* size of the host block                                                  */
    int32_t record_words;   /* obs_dim + 2                                                             */
    int32_t prefix_records; /* records that travel with the first copy (the block holds this many)     */
} os2r_packed_layout;
int32_t os2r_packed_layout_get(const os2r_env *env, int32_t prefix_records, os2r_packed_layout *out);
/* actions[N,2] (host) -> block (host, laid out as above; page-locked memory is written by DMA directly, pageable
 * memory through the handle's staging block). *n_terminal = number of finished envs; when it exceeds
 * prefix_records the remaining records are fetched with os2r_fetch_terminal_records before the next step. */
int32_t os2r_step_host_packed(os2r_env *env, const float *actions, void *block, int32_t prefix_records,
                              int32_t *n_terminal);
/* The same step in two halves, for VecEnv.step_async / step_wait (subproc_vec_env.py:114-123: send the actions, do
 * other work, collect): _begin stages the actions and enqueues H2D + kernel + D2H on the handle's stream and returns;
 * _end waits and completes the block. `actions` may be reused after _begin returns; `block` must stay untouched
 * until _end. One step in flight per handle. */
int32_t os2r_step_host_packed_begin(os2r_env *env, const float *actions, void *block, int32_t prefix_records);
int32_t os2r_step_host_packed_end(os2r_env *env, int32_t *n_terminal);
int32_t os2r_fetch_terminal_records(os2r_env *env, int32_t first, int32_t count, int32_t *records_host);

/* Packed double state [N, os2r_state_width] <-> device SoA (checkpoint/resume + parity tests). */
int32_t os2r_get_state(os2r_env *env, double *state_host);
int32_t os2r_set_state(os2r_env *env, const double *state_host);
int32_t os2r_get_params(os2r_env *env, double *params_host);
int32_t os2r_set_params(os2r_env *env, const double *params_host);
/* Per-env episode bookkeeping: steps[N] (int32), returns[N] (double), reset_ids[N] (int32 index
 * into the task's reset_positions = info['reset_orientation'], tasks/monopod.py:374). Any may be NULL. */
int32_t os2r_get_episode(os2r_env *env, int32_t *steps_host, double *returns_host, int32_t *reset_ids_host,
                         uint32_t *episodes_host);
/* Restore the bookkeeping (checkpoint / resume into a NEW handle): `episodes` is the per-env episode counter = the
 * counter word of the env's RNG stream, so a restored env draws the same reset poses / parameters it would have
 * drawn without the interruption. Any pointer may be NULL (left unchanged). */
int32_t os2r_set_episode(os2r_env *env, const int32_t *steps_host, const double *returns_host,
                         const int32_t *reset_ids_host, const uint32_t *episodes_host);

int32_t os2r_stats_read(os2r_env *env, os2r_stats *out, int32_t clear);
int32_t os2r_stats_write(os2r_env *env, const os2r_stats *in); /* checkpoint / resume */

/* Page-locked host array [N,2] float32 owned by the handle: a caller that writes its actions HERE and passes this
 * pointer to os2r_step_host / os2r_step_host_packed(_begin) skips the staging memcpy (the H2D copy reads it directly). */
int32_t os2r_host_action_buffer(os2r_env *env, float **actions_out);

/* Introspection used by bench/tests. */
int64_t os2r_num_envs(const os2r_env *env);
int32_t os2r_obs_dim(const os2r_env *env);
int64_t os2r_kernel_launches(const os2r_env *env); /* kernels launched by this handle so far */
/* Launch geometry and static resource use of the step kernel (for the roofline report);
 * resident_blocks_per_sm comes from the CUDA occupancy API. Any out pointer may be NULL. */
int32_t os2r_kernel_info(const os2r_env *env, int32_t *block_threads, int32_t *grid_blocks,
                         int32_t *regs_per_thread, int32_t *local_bytes_per_thread,
                         int32_t *resident_blocks_per_sm, int32_t *envs_per_thread);
/* Structure signature of a model (which constant-table entries are exactly 0 / 1; gym_os2r_b200/csrc/os2r_device.cuh)
 * and whether this library carries step kernels specialised on it (the shipped URDFs: yes; any other model runs the
 * all-general kernels). Host only: needs no GPU. Any out pointer may be NULL. */
int32_t os2r_model_signature(const os2r_model *model, uint32_t *joints, uint32_t *contacts, int32_t *specialised);
/* Violation counters of the CHECKED build (`make -C gym_os2r_b200/csrc libos2r_checked.so`: the step kernel validates
 * its lane-sort permutation, env indices, terminal-record capacity and a shared-memory guard word; violations are
 * counted, not trapped). out8[0..4] = counts (see os2r_kernels.cu), out8[7] = 1 when the loaded library was built
 * with the checks (0: the counters are always zero). */
int32_t os2r_debug_counters(int32_t device, uint64_t *out8, int32_t clear);
/* fp32 FMA-pipe peak microbenchmark on the handle's device: returns TFLOP/s (2 flop per FMA)
 * measured with CUDA events (MEASURED_PEAKS.json has no fp32 entry). */
int32_t os2r_measure_fp32_peak(int32_t device, double *tflops_out, double *sm_clock_mhz_out);

#ifdef __cplusplus
}
#endif
#endif /* OS2R_H */
