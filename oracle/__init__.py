"""ctypes binding of the CPU oracle (oracle/os2r_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs (and by the diagnostic scripts under tools/ that are themselves test infrastructure:
parity report, regression-fixture generator, probes). The product package never imports this module.
Physics parity is UNPINNED (no runnable DART here, no golden trajectories in the reference);
task-logic parity is pinned by tests/golden/task_kat.json. See the header of os2r_oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from gym_os2r_b200 import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_build', 'libos2r_oracle.so')
_lib = None

SIGMOID_IDS = {name: i for i, name in enumerate(
    ('gaussian', 'hyperbolic', 'long_tail', 'reciprocal', 'cosine', 'linear', 'quadratic', 'tanh_squared'))}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, 'os2r_oracle.c')
    hdr = os.path.join(_HERE, '..', 'include', 'os2r.h')
    stale = (not os.path.exists(LIB_PATH)
             or (os.path.exists(src) and os.path.getmtime(LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))))
    if force or stale:
        subprocess.check_call(['make', '-C', _HERE, '-B', '_build/libos2r_oracle.so'], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_uniform.restype = C.c_double
        _lib.oracle_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        _lib.oracle_tolerance.restype = C.c_double
        _lib.oracle_tolerance.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double]
        _lib.oracle_energy.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _d(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Oracle:
    """N monopod envs stepped on the CPU in fp64 (same state / params packing as the C-ABI)."""

    def __init__(self, model: _capi.Model, task: _capi.TaskCfg, n_envs: int, first_env_id: int = 0,
                 seed: int = 0, nthreads: int = 1):
        self.L = lib()
        self.model, self.task = model, task
        self.N, self.first, self.seed_value, self.nthreads = int(n_envs), int(first_env_id), int(seed), int(nthreads)
        self.W = _capi.state_width(model)
        self.PW = _capi.params_width(model)
        self.D = task.obs_dim
        self.state = np.zeros((self.N, self.W))
        self.params = np.zeros((self.N, self.PW))
        self.steps = np.zeros(self.N, dtype=np.int32)
        self.ret = np.zeros(self.N)
        self.episode = np.zeros(self.N, dtype=np.uint32)
        self.reset_id = np.zeros(self.N, dtype=np.int32)
        self.L.oracle_init(C.byref(model), C.byref(task), C.c_int64(self.N), C.c_int64(self.first),
                           C.c_uint64(self.seed_value), _p(self.state), _p(self.params), _p(self.steps),
                           _p(self.ret), _p(self.episode), _p(self.reset_id))

    def _common(self):
        return (C.byref(self.model), C.byref(self.task), C.c_int64(self.N), C.c_int64(self.first),
                C.c_uint64(self.seed_value), _p(self.state), _p(self.params), _p(self.steps), _p(self.ret),
                _p(self.episode), _p(self.reset_id))

    def seed(self, seed: int):
        self.seed_value = int(seed)

    def reset(self, mask=None):
        obs = np.zeros((self.N, self.D))
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.oracle_reset(*self._common(), _p(m), _p(obs))
        return obs

    def step(self, actions):
        a = _d(actions, (self.N, 2))
        obs = np.zeros((self.N, self.D))
        rew = np.zeros(self.N)
        done = np.zeros(self.N, dtype=np.uint8)
        term = np.zeros((self.N, self.D))
        info = np.zeros((self.N, 2), dtype=np.int32)
        self.L.oracle_step(*self._common(), _p(a), _p(obs), _p(rew), _p(done), _p(term), _p(info),
                           C.c_int32(self.nthreads))
        return obs, rew, done.astype(bool), term, info

    # views into the packed state
    @property
    def n(self):
        return self.model.n_dof

    @property
    def q(self):
        return self.state[:, :self.n]

    @property
    def qd(self):
        return self.state[:, self.n:2 * self.n]


# ---- single-state hooks -------------------------------------------------------------------------

def nominal_params(model: _capi.Model) -> np.ndarray:
    n, nc = model.n_dof, model.n_contacts
    row = np.zeros(_capi.params_width(model))
    row[:n] = 1.0
    row[n:2 * n] = model.damping[:n]
    row[2 * n:3 * n] = model.friction[:n]
    row[3 * n:3 * n + nc] = model.contact_mu[:nc]
    row[3 * n + nc] = model.gravity_z
    return row


def dynamics_debug(model, params_row, q, v, action):
    n, nc = model.n_dof, model.n_contacts
    rows = n + 3 * nc
    qdd, Minv, depth, J = np.zeros(n), np.zeros((n, n)), np.zeros(max(nc, 1)), np.zeros((rows, n))
    lib().oracle_dynamics_debug(C.byref(model), _p(_d(params_row)), _p(_d(q)), _p(_d(v)), _p(_d(action)),
                                _p(qdd), _p(Minv), _p(depth), _p(J))
    return qdd, Minv, depth[:nc], J


def forward_dynamics_plain(model, params_row, q, v, tau):
    qdd = np.zeros(model.n_dof)
    lib().oracle_forward_dynamics_plain(C.byref(model), _p(_d(params_row)), _p(_d(q)), _p(_d(v)), _p(_d(tau)), _p(qdd))
    return qdd


def substeps(model, params_row, q, v, lam, action, count, sweeps=0, tol=0.0):
    q, v, lam = _d(q).copy(), _d(v).copy(), _d(lam).copy()
    lib().oracle_substeps(C.byref(model), _p(_d(params_row)), _p(q), _p(v), _p(lam), _p(_d(action)),
                          C.c_int(count), C.c_int(sweeps), C.c_double(tol))
    return q, v, lam


def evaluate(model, task, q, v, a0, a1):
    D = task.obs_dim
    raw, obs, rew, done = np.zeros(D), np.zeros(D), C.c_double(0), C.c_int32(0)
    lib().oracle_evaluate(C.byref(model), C.byref(task), _p(_d(q)), _p(_d(v)), _p(_d(a0)), _p(_d(a1)),
                          _p(raw), _p(obs), C.byref(rew), C.byref(done))
    return raw, obs, rew.value, bool(done.value)


def state_info(task, obs, a0, a1):
    rew, done = C.c_double(0), C.c_int32(0)
    lib().oracle_state_info(C.byref(task), _p(_d(obs)), _p(_d(a0)), _p(_d(a1)), C.byref(rew), C.byref(done))
    return rew.value, bool(done.value)


def tolerance(x, bounds=(0.0, 0.0), margin=0.0, sigmoid='gaussian', value_at_margin=0.1):
    return lib().oracle_tolerance(float(x), float(bounds[0]), float(bounds[1]), float(margin),
                                  SIGMOID_IDS[sigmoid], float(value_at_margin))


def leg_joint_angles(task, pitch):
    out = np.zeros(2)
    lib().oracle_leg_joint_angles(C.byref(task), C.c_double(pitch), _p(out))
    return out


def fk(model, q):
    bp, cc = np.zeros((model.n_dof, 3)), np.zeros((max(model.n_contacts, 1), 3))
    lib().oracle_fk(C.byref(model), _p(_d(q)), _p(bp), _p(cc))
    return bp, cc[:model.n_contacts]


def energy(model, params_row, q, v):
    return lib().oracle_energy(C.byref(model), _p(_d(params_row)), _p(_d(q)), _p(_d(v)))


def uniform(seed, env_id, episode, k):
    return lib().oracle_uniform(seed, env_id, episode, k)
