"""ctypes binding of the C-ABI in ``include/os2r.h`` (libos2r.so).

Only plain pointers and sizes cross this boundary; torch is used by callers purely to own device
buffers and streams. There is no CPU fallback: if the CUDA library cannot be loaded, or a handle
cannot be created on a CUDA device, the error is raised to the caller.
"""
import ctypes as C
import os

MAX_DOF = 5
MAX_CONTACTS = 6
MAX_OBS = 12
MAX_RESETS = 8
MAX_ROWS = MAX_DOF + 3 * MAX_CONTACTS
N_ROLES = 5
ABI_VERSION = 9

ROLE_HIP, ROLE_KNEE, ROLE_PITCH, ROLE_YAW, ROLE_BOOM_CONNECTOR = range(5)
ROLE_OF_JOINT = {
    'hip_joint': ROLE_HIP,
    'knee_joint': ROLE_KNEE,
    'planarizer_pitch_joint': ROLE_PITCH,
    'planarizer_yaw_joint': ROLE_YAW,
    'boom_connector_joint': ROLE_BOOM_CONNECTOR,
}
OBS_POS, OBS_POS_PERIODIC, OBS_VEL, OBS_TORQUE = range(4)
(REWARD_CUSTOM, REWARD_BALANCING_V1, REWARD_BALANCING_V2, REWARD_BALANCING_V3,
 REWARD_HOPPING_V1, REWARD_STRAIGHT_V1) = range(6)

_i32 = C.c_int32
_f64 = C.c_double


class Model(C.Structure):
    """struct os2r_model"""
    _fields_ = [
        ('n_dof', _i32), ('n_contacts', _i32), ('substeps', _i32), ('pgs_iters', _i32),
        ('axis', _i32 * MAX_DOF), ('role_dof', _i32 * N_ROLES),
        ('contact_body', _i32 * MAX_CONTACTS), ('pgs_joint_sweeps', _i32),
        ('tree_R', (_f64 * 9) * MAX_DOF), ('tree_p', (_f64 * 3) * MAX_DOF),
        ('mass', _f64 * MAX_DOF), ('com', (_f64 * 3) * MAX_DOF),
        ('inertia', (_f64 * 6) * MAX_DOF),
        ('damping', _f64 * MAX_DOF), ('friction', _f64 * MAX_DOF),
        ('contact_pos', (_f64 * 3) * MAX_CONTACTS), ('contact_radius', _f64 * MAX_CONTACTS),
        ('contact_mu', _f64 * MAX_CONTACTS),
        ('gravity_z', _f64), ('dt', _f64), ('erp', _f64), ('max_erv', _f64),
        ('cfm_contact', _f64), ('cfm_joint', _f64), ('max_torque', _f64 * 2), ('pgs_tol', _f64),
    ]


class TaskCfg(C.Structure):
    """struct os2r_task_cfg"""
    _fields_ = [
        ('obs_dim', _i32), ('normalized', _i32), ('reward_id', _i32), ('max_episode_steps', _i32),
        ('auto_reset', _i32), ('n_resets', _i32), ('reset_randomized', _i32),
        ('randomize_params', _i32), ('randomize_gravity', _i32), ('simple_sample_reset', _i32),
        ('gravity_redraw_resets', _i32), ('_pad1', _i32),
        ('reward_pitch_col', _i32), ('reward_yawvel_col', _i32), ('reward_hip_col', _i32),
        ('reward_knee_col', _i32),
        ('obs_kind', _i32 * MAX_OBS), ('obs_index', _i32 * MAX_OBS),
        ('reset_laying', _i32 * MAX_RESETS),
        ('obs_low', _f64 * MAX_OBS), ('obs_high', _f64 * MAX_OBS),
        ('done_low', _f64 * MAX_OBS), ('done_high', _f64 * MAX_OBS),
        ('reset_pitch', _f64 * MAX_RESETS),
        ('simple_lo', _f64 * 2), ('simple_hi', _f64 * 2),
        ('ik_upper_leg', _f64), ('ik_lower_leg', _f64), ('ik_pivot_height', _f64),
        ('ik_boom', _f64), ('ik_hip_offset', _f64), ('ik_clip', _f64),
        ('mass_lo', _f64), ('mass_hi', _f64), ('fric_lo', _f64), ('fric_hi', _f64),
        ('damp_lo', _f64), ('damp_hi', _f64), ('mu_lo', _f64), ('mu_hi', _f64), ('mu_link', _f64),
        ('grav_mean', _f64), ('grav_std', _f64),
    ]


class Tuning(C.Structure):
    """struct os2r_tuning (zero = defaults)"""
    _fields_ = [('sort_margin', _f64), ('force_block', _i32), ('disable_specialisation', _i32), ('disable_root_fold', _i32),
                ('_pad', _i32)]


class PackedLayout(C.Structure):
    """struct os2r_packed_layout"""
    _fields_ = [
        ('obs', C.c_int64), ('reward', C.c_int64), ('done', C.c_int64), ('reset_id', C.c_int64),
        ('term_count', C.c_int64), ('term_records', C.c_int64), ('total_bytes', C.c_int64),
        ('record_words', _i32), ('prefix_records', _i32),
    ]


class Stats(C.Structure):
    """struct os2r_stats"""
    _fields_ = [
        ('env_steps', C.c_uint64), ('episodes', C.c_uint64), ('done_task', C.c_uint64),
        ('done_timelimit', C.c_uint64), ('nonfinite_resets', C.c_uint64),
        ('sum_return', _f64), ('sum_length', _f64),
    ]


def state_width(model: Model) -> int:
    return 2 * model.n_dof + (model.n_dof + 3 * model.n_contacts) + 2


def params_width(model: Model) -> int:
    return 3 * model.n_dof + model.n_contacts + 1


# every symbol include/os2r.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
SYMBOLS = {
    'os2r_abi_version': (_i32, []),
    'os2r_last_error': (C.c_char_p, []),
    'os2r_state_width': (_i32, [C.POINTER(Model)]),
    'os2r_params_width': (_i32, [C.POINTER(Model)]),
    'os2r_create': (_i32, [C.POINTER(Model), C.POINTER(TaskCfg), C.c_int64, C.c_int64, _i32,
                           C.c_uint64, _i32, C.POINTER(_vp)]),
    'os2r_create_tuned': (_i32, [C.POINTER(Model), C.POINTER(TaskCfg), C.c_int64, C.c_int64, _i32,
                                 C.c_uint64, _i32, C.POINTER(Tuning), C.POINTER(_vp)]),
    'os2r_destroy': (_i32, [_vp]),
    'os2r_set_randomization': (_i32, [_vp, C.POINTER(TaskCfg)]),
    'os2r_seed': (_i32, [_vp, C.c_uint64]),
    'os2r_reset': (_i32, [_vp, _vp, _vp, _vp]),
    'os2r_step': (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'os2r_step_host': (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'os2r_packed_layout_get': (_i32, [_vp, _i32, C.POINTER(PackedLayout)]),
    'os2r_step_host_packed': (_i32, [_vp, _vp, _vp, _i32, C.POINTER(_i32)]),
    'os2r_step_host_packed_begin': (_i32, [_vp, _vp, _vp, _i32]),
    'os2r_step_host_packed_end': (_i32, [_vp, C.POINTER(_i32)]),
    'os2r_fetch_terminal_records': (_i32, [_vp, _i32, _i32, _vp]),
    'os2r_get_state': (_i32, [_vp, _vp]),
    'os2r_set_state': (_i32, [_vp, _vp]),
    'os2r_get_params': (_i32, [_vp, _vp]),
    'os2r_set_params': (_i32, [_vp, _vp]),
    'os2r_get_episode': (_i32, [_vp, _vp, _vp, _vp, _vp]),
    'os2r_set_episode': (_i32, [_vp, _vp, _vp, _vp, _vp]),
    'os2r_stats_read': (_i32, [_vp, C.POINTER(Stats), _i32]),
    'os2r_stats_write': (_i32, [_vp, C.POINTER(Stats)]),
    'os2r_host_action_buffer': (_i32, [_vp, C.POINTER(C.POINTER(C.c_float))]),
    'os2r_num_envs': (C.c_int64, [_vp]),
    'os2r_obs_dim': (_i32, [_vp]),
    'os2r_kernel_launches': (C.c_int64, [_vp]),
    'os2r_kernel_info': (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32),
                                C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    'os2r_model_signature': (_i32, [C.POINTER(Model), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(_i32)]),
    'os2r_debug_counters': (_i32, [_i32, C.POINTER(C.c_uint64), _i32]),
    'os2r_measure_fp32_peak': (_i32, [_i32, C.POINTER(_f64), C.POINTER(_f64)]),
}

LIB_PATH = os.environ.get('OS2R_LIB') or os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc', 'libos2r.so')
_lib = None


class Os2rError(RuntimeError):
    """Raised when a C-ABI call returns non-zero (message from os2r_last_error)."""


def load_library(path: str = None):
    """Load libos2r.so and type every exported entry point. Fails loudly when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise Os2rError(
            f'{p} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(or `make -C gym_os2r_b200/csrc`). There is no CPU fallback for the step path.')
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    got = lib.os2r_abi_version()
    if got != ABI_VERSION:
        raise Os2rError(f'libos2r.so ABI version {got} != binding version {ABI_VERSION}; rebuild')
    if path is None:
        _lib = lib
    return lib


def check(status: int, lib=None):
    if status != 0:
        lib = lib or load_library()
        msg = lib.os2r_last_error()
        raise Os2rError(msg.decode() if msg else f'os2r call failed with status {status}')
