"""CPU suite: cross-checks that pin the physics half of the oracle WITHOUT DART:
an independent Lagrangian mass matrix, finite-difference Jacobians, energy conservation, static
equilibrium, and the fixed-sweep PGS against the same sweeps run to convergence."""
import numpy as np
import pytest

import oracle
from gym_os2r_b200.models import compiler

from helpers import make_config


def _mass_matrix_numpy(cm, params, q):
    """M = sum_b m Jv^T Jv + Jw^T (R I R^T) Jw from numpy forward kinematics (independent of the C code)."""
    m = cm.struct
    n = m.n_dof
    Rs, ps, _ = compiler.forward_kinematics(m, q)
    axes = []
    R, p = np.eye(3), np.zeros(3)
    for i in range(n):
        Rt = np.array(m.tree_R[i][:]).reshape(3, 3)
        A = (Rs[i - 1] if i else np.eye(3)) @ Rt
        axes.append(A[:, m.axis[i]])
    M = np.zeros((n, n))
    for b in range(n):
        mass = m.mass[b] * params[b]
        c = ps[b] + Rs[b] @ np.array(m.com[b][:])
        t = m.inertia[b]
        Ib = np.array([[t[0], t[3], t[4]], [t[3], t[1], t[5]], [t[4], t[5], t[2]]])
        Iw = Rs[b] @ Ib @ Rs[b].T
        Jv, Jw = np.zeros((3, n)), np.zeros((3, n))
        for j in range(b + 1):
            Jv[:, j] = np.cross(axes[j], c - ps[j])
            Jw[:, j] = axes[j]
        M += mass * Jv.T @ Jv + Jw.T @ Iw @ Jw
    return M


@pytest.mark.parametrize('mode', ['simple', 'fixed', 'fixed_hip', 'free_hip'])
def test_minv_matches_independent_mass_matrix(mode):
    task, cm, cfg = make_config(mode, reward='StraightV1' if mode == 'simple' else 'BalancingV1')
    m = cm.struct
    rng = np.random.RandomState(0)
    p = oracle.nominal_params(m)
    p[:m.n_dof] = rng.uniform(0.8, 1.2, m.n_dof)
    for _ in range(5):
        q = rng.uniform(-2, 2, m.n_dof)
        _, Minv, _, _ = oracle.dynamics_debug(m, p, q, np.zeros(m.n_dof), [0, 0])
        M = _mass_matrix_numpy(cm, p, q)
        np.testing.assert_allclose(Minv @ M, np.eye(m.n_dof), atol=1e-9)
        assert np.abs(Minv - Minv.T).max() < 1e-9 and np.linalg.eigvalsh(M).min() > 0


def test_forward_dynamics_satisfies_equation_of_motion():
    """M qdd + h = tau with h from the oracle's own zero-torque solve: linearity in tau (ABA is consistent)."""
    task, cm, cfg = make_config('free_hip', reward='BalancingV1')
    m = cm.struct
    n = m.n_dof
    rng = np.random.RandomState(1)
    p = oracle.nominal_params(m)
    q, v = rng.uniform(-1, 1, n), rng.uniform(-2, 2, n)
    M = _mass_matrix_numpy(cm, p, q)
    a0 = oracle.forward_dynamics_plain(m, p, q, v, np.zeros(n))
    for _ in range(4):
        tau = rng.uniform(-3, 3, n)
        a = oracle.forward_dynamics_plain(m, p, q, v, tau)
        np.testing.assert_allclose(M @ (a - a0), tau, atol=1e-9)


def test_contact_jacobian_matches_finite_differences():
    task, cm, cfg = make_config('fixed_hip', reward='BalancingV1')
    m = cm.struct
    n = m.n_dof
    q = np.array([0.1, -0.03, 1.0, -2.2])       # lying: contacts penetrate so rows are built
    _, _, depth, J = oracle.dynamics_debug(m, oracle.nominal_params(m), q, np.zeros(n), [0, 0])
    assert (depth > 0).any()
    eps = 1e-6
    for c in range(m.n_contacts):
        if depth[c] <= 0:
            assert np.all(J[n + 3 * c:n + 3 * c + 3] == 0)
            continue
        Rs, ps, _ = compiler.forward_kinematics(m, q)
        for i in range(n):
            dq = np.zeros(n); dq[i] = eps
            _, cp = oracle.fk(m, q + dq)
            _, cm_ = oracle.fk(m, q - dq)
            d = (cp[c] - cm_[c]) / (2 * eps)            # velocity of the sphere CENTRE per unit joint rate
            # the constraint acts at the lowest point of the sphere (centre - r z): add axis x (-r z)
            Rt = np.array(m.tree_R[i][:]).reshape(3, 3)
            axis = ((Rs[i - 1] if i else np.eye(3)) @ Rt)[:, m.axis[i]]
            if i <= m.contact_body[c]:
                d = d + np.cross(axis, [0, 0, -m.contact_radius[c]])
            np.testing.assert_allclose(J[n + 3 * c:n + 3 * c + 3, i], [d[2], d[0], d[1]], atol=1e-7)


def test_energy_conservation_rk4():
    """Undamped, frictionless, torque-free, contact-free motion integrated with RK4 at dt = 1e-4 conserves energy."""
    task, cm, cfg = make_config('free_hip', reward='BalancingV1')
    m = cm.struct
    n = m.n_dof
    p = oracle.nominal_params(m)
    p[n:3 * n] = 0.0
    rng = np.random.RandomState(0)
    x = np.concatenate([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)])
    x[1] = 0.6
    f = lambda s: np.concatenate([s[n:], oracle.forward_dynamics_plain(m, p, s[:n], s[n:], np.zeros(n))])
    e0 = oracle.energy(m, p, x[:n], x[n:])
    h = 1e-4
    for _ in range(600):
        k1 = f(x); k2 = f(x + h / 2 * k1); k3 = f(x + h / 2 * k2); k4 = f(x + h * k3)
        x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    assert abs(oracle.energy(m, p, x[:n], x[n:]) - e0) < 1e-11 * max(1.0, abs(e0))


def test_static_stand_normal_force():
    """Monopod resting on its foot with locked-ish leg: total normal impulse per iteration = supported weight * dt
    (moment balance about the pitch axis gives ~7 N at the foot; SURVEY.md section 8c cross-check iv)."""
    task, cm, cfg = make_config('fixed_hip', reward='BalancingV1')
    m = cm.struct
    n = m.n_dof
    p = oracle.nominal_params(m)
    p[2 * n:3 * n] = 0.0     # strong friction on hip and knee only: the leg holds its pose, the boom is free
    p[2 * n + cm.dof_of('hip_joint')] = p[2 * n + cm.dof_of('knee_joint')] = 5.0
    q = np.zeros(n)
    q[cm.dof_of('planarizer_pitch_joint')], q[cm.dof_of('hip_joint')], q[cm.dof_of('knee_joint')] = 0.12, 0.6375366, -1.3146489
    v, lam = np.zeros(n), np.zeros(n + 9)
    q, v, lam = oracle.substeps(m, p, q, v, lam, [0, 0], 4000)
    foot_force = lam[n + 6] / m.dt
    assert 6.5 < foot_force < 7.7, foot_force
    assert np.abs(v).max() < 1e-2


def test_fixed_sweeps_track_converged_lcp():
    """8 warm-started sweeps per iteration stay close to the same PGS run to convergence."""
    task, cm, cfg = make_config('fixed_hip', reward='BalancingV1')
    m = cm.struct
    n = m.n_dof
    p = oracle.nominal_params(m)
    q0 = np.zeros(n)
    q0[cm.dof_of('planarizer_pitch_joint')], q0[cm.dof_of('hip_joint')], q0[cm.dof_of('knee_joint')] = 0.15, 0.2861, -0.5877
    rng = np.random.RandomState(5)
    acts = 0.1 * rng.uniform(-1, 1, (120, 2))

    def roll(sweeps, tol, joint_sweeps=1):
        m.pgs_joint_sweeps = joint_sweeps
        q, v, lam = q0.copy(), np.zeros(n), np.zeros(n + 9)
        for a in acts:
            q, v, lam = oracle.substeps(m, p, q, v, lam, a, 10, sweeps=sweeps, tol=tol)
        return q
    ref = roll(4000, 1e-16, joint_sweeps=0)             # every row in every sweep, to convergence
    assert np.abs(roll(8, 0.0, joint_sweeps=0) - ref).max() < 2e-3
    assert np.abs(roll(8, 0.0) - ref).max() < 2e-3      # ~30 steps after touchdown
    assert np.abs(roll(8, 0.0) - ref).max() < np.abs(roll(2, 0.0) - ref).max()


def test_sweep_tolerance_rule():
    """The per-env sweep exit (os2r_model.pgs_tol): a sweep that moved the velocity by <= tol in the kinetic-energy
    norm ends the iteration. Checked on the oracle: (i) contact-free, the production tolerance changes nothing at all
    (saturated / sticking joint friction converges in the first sweep); (ii) in contact, it costs no accuracy against
    fully converged sweeps compared with always running the cap; (iii) it removes most sweeps; (iv) the measure is
    the energy norm: the inverse the rule uses reproduces the independent numpy mass matrix."""
    import ctypes as C
    lib = oracle.lib()
    hist = (C.c_longlong * 65)()

    def sweeps_done():
        lib.oracle_sweep_histogram(hist, 1)
        h = np.array(hist[:])
        return (h * np.arange(65)).sum(), h.sum()

    # (i) contact-free: `simple` never touches the ground
    task, cm, cfg = make_config('simple', reward='StraightV1')
    m, n = cm.struct, cm.n_dof
    p = oracle.nominal_params(m)
    rng = np.random.RandomState(2)
    q, v, lam = rng.uniform(-1, 1, n), rng.uniform(-2, 2, n), np.zeros(n + 9)
    a = np.array([0.3, -0.2])
    out_fixed = oracle.substeps(m, p, q, v, lam, a, 500, sweeps=8, tol=0.0)
    sweeps_done()
    out_tol = oracle.substeps(m, p, q, v, lam, a, 500, sweeps=8, tol=1e-6)
    total, iters = sweeps_done()
    assert np.array_equal(out_fixed[0], out_tol[0]) and np.array_equal(out_fixed[1], out_tol[1])
    assert iters == 500 and total <= 2 * iters

    # (ii)/(iii) with contact: dropped fixed_hip monopod, 120 env steps (touchdown at ~90)
    task, cm, cfg = make_config('fixed_hip', reward='BalancingV1')
    m, n = cm.struct, cm.n_dof
    p = oracle.nominal_params(m)
    q0 = np.zeros(n)
    q0[cm.dof_of('planarizer_pitch_joint')], q0[cm.dof_of('hip_joint')], q0[cm.dof_of('knee_joint')] = 0.15, 0.2861, -0.5877
    acts = 0.1 * np.random.RandomState(5).uniform(-1, 1, (120, 2))

    def roll(sweeps, tol, joint_sweeps=1):
        m.pgs_joint_sweeps = joint_sweeps
        q, v, lam = q0.copy(), np.zeros(n), np.zeros(n + 9)
        for a in acts:
            q, v, lam = oracle.substeps(m, p, q, v, lam, a, 10, sweeps=sweeps, tol=tol)
        return q
    ref = roll(4000, 1e-16, joint_sweeps=0)             # every row in every sweep, to convergence
    sweeps_done()
    e_fixed = np.abs(roll(8, 0.0) - ref).max()
    total_fixed, iters = sweeps_done()
    e_tol = np.abs(roll(8, 1e-6) - ref).max()
    total_tol, iters2 = sweeps_done()
    assert iters == iters2 == 1200 and total_fixed <= 8 * iters
    assert e_tol < 2e-3 and e_tol < 2 * e_fixed + 1e-6, (e_tol, e_fixed)
    assert total_tol < total_fixed, (total_tol, total_fixed)   # tol 0 already stops at exact fixed points (no contact)

    # (iv) Minv @ M_numpy = I: the matrix the rule inverts is the inverse of the true mass matrix
    qq = rng.uniform(-1, 1, n)
    _, Minv, _, _ = oracle.dynamics_debug(m, p, qq, np.zeros(n), np.zeros(2))
    np.testing.assert_allclose(Minv @ _mass_matrix_numpy(cm, p, qq), np.eye(n), atol=1e-9)


def test_converged_sweeps_solve_the_boxed_lcp():
    """Solver-independent check of the constraint stage: run the oracle's sweeps to convergence from random
    penetrating states and verify the KKT conditions of the boxed LCP DART poses (the problem its Dantzig solver
    answers exactly) with matrices rebuilt here in numpy — w = J (v* + Minv J^T lam) - target + cfm A_rr lam_r must be
    >= 0 where lam sits on its lower bound, <= 0 on its upper bound, and 0 in between; bounds: joint friction
    +-f dt, normal [0, inf), tangents +-mu lam_n."""
    task, cm, cfg = make_config('fixed_hip', reward='BalancingV1', pgs_joint_sweeps=0)   # every row in every sweep
    m, n, nc = cm.struct, cm.n_dof, cm.struct.n_contacts
    p = oracle.nominal_params(m)
    p[2 * n:3 * n] = [0.02, 0.03, 0.015, 0.04]          # joint friction N m
    mu = p[3 * n:3 * n + nc]
    rng = np.random.RandomState(8)
    checked_contacts = 0
    for trial in range(12):
        q = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.05, -0.02), rng.uniform(0.6, 1.4), rng.uniform(-2.6, -1.8)])
        v = rng.normal(0, 1.0, n)
        a = rng.uniform(-1, 1, 2)
        qdd, Minv, depth, J = oracle.dynamics_debug(m, p, q, v, a)
        act = np.repeat(depth > 0, 3)
        rows = np.concatenate([np.ones(n, bool), act])
        vstar = v + m.dt * qdd
        # one physics iteration with (practically) converged sweeps, cold start
        q1, v1, lam = oracle.substeps(m, p, q, v, np.zeros(n + 3 * nc), a, 1, sweeps=20000, tol=1e-18)
        np.testing.assert_allclose(v1, vstar + Minv @ J.T @ lam, atol=1e-12)        # impulses act through plain Minv
        np.testing.assert_allclose(q1, q + m.dt * v1, atol=1e-15)                   # semi-implicit position update
        A = J @ Minv @ J.T
        cfm = np.concatenate([np.full(n, m.cfm_joint), np.full(3 * nc, m.cfm_contact)])
        target = np.zeros(n + 3 * nc)
        target[n::3] = np.minimum(np.maximum(depth, 0) * m.erp / m.dt, m.max_erv)
        w = J @ v1 - target + cfm * np.diag(A) * lam
        lo, hi = np.zeros_like(lam), np.zeros_like(lam)
        hi[:n] = p[2 * n:3 * n] * m.dt; lo[:n] = -hi[:n]
        for c in range(nc):
            lo[n + 3 * c], hi[n + 3 * c] = 0.0, np.inf
            hi[n + 3 * c + 1:n + 3 * c + 3] = mu[c] * lam[n + 3 * c]
            lo[n + 3 * c + 1:n + 3 * c + 3] = -mu[c] * lam[n + 3 * c]
        scale = np.abs(J @ vstar).max() + 1e-3
        for r in np.flatnonzero(rows):
            at_lo, at_hi = lam[r] <= lo[r] + 1e-15, lam[r] >= hi[r] - 1e-15
            if at_lo and at_hi:
                continue                                    # degenerate box (mu * 0): any w is admissible
            if at_lo:
                assert w[r] >= -1e-9 * scale, (trial, r, w[r])
            elif at_hi:
                assert w[r] <= 1e-9 * scale, (trial, r, w[r])
            else:
                assert abs(w[r]) <= 1e-9 * scale, (trial, r, w[r], lam[r], lo[r], hi[r])
        assert np.all(lam[~rows] == 0)
        checked_contacts += int((lam[n::3] > 0).sum())
    assert checked_contacts >= 10
