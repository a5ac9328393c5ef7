"""CPU suite: the oracle (oracle/os2r_oracle.c) and the product's host logic against the golden
vectors recorded from the reference's own numpy code (tools/gen_golden.py)."""
import numpy as np
import pytest

import oracle
from gym_os2r_b200 import rewards
from gym_os2r_b200.rewards import rewards_utils
from gym_os2r_b200.utils.reset import leg_joint_angles

from helpers import chain_state, make_config


def test_spaces_match_reference(golden):
    for sp in golden['spaces']:
        reward = 'StraightV1' if sp['task_mode'] == 'simple' else 'BalancingV1'
        task, compiled, cfg = make_config(sp['task_mode'], sp['variant'], reward)
        assert task.joint_names == sp['joint_names']
        assert task.action_names == sp['action_names']
        assert task.observation_index == sp['observation_index']
        assert list(task.observation_mask) == sp['observation_mask']
        assert list(task.periodic_joints) == sp['periodic_joints']
        np.testing.assert_array_equal(task.observation_space.low, sp['obs_low'])
        np.testing.assert_array_equal(task.observation_space.high, sp['obs_high'])
        np.testing.assert_array_equal(task.reset_space.low, sp['reset_low'])
        np.testing.assert_array_equal(task.reset_space.high, sp['reset_high'])
        np.testing.assert_array_equal(task.max_torques, sp['max_torques'])
        assert cfg.obs_dim == len(sp['obs_low'])


def _configs(golden):
    keys = sorted({(k['task_mode'], k['variant'], k['reward']) for k in golden['task_kat']})
    return keys


def test_task_kat_oracle_and_host(golden):
    """obs bit-exact... to 1 ulp (libm tanh), reward exact, done exact — for the C oracle, for the
    host formulas of the product Task, and for the raw-threshold done rule the device uses."""
    by_cfg = {}
    for k in golden['task_kat']:
        by_cfg.setdefault((k['task_mode'], k['variant'], k['reward']), []).append(k)
    assert len(by_cfg) >= 40
    n_done = 0
    for (mode, variant, reward), cases in by_cfg.items():
        task, compiled, cfg = make_config(mode, variant, reward)
        m = compiled.struct
        for k in cases:
            q, v = chain_state(task, compiled, k['q'], k['v'])
            raw, obs, rew, done = oracle.evaluate(m, cfg, q, v, k['a0'], k['a1'])
            np.testing.assert_allclose(obs, k['obs'], rtol=0, atol=5e-16, err_msg=f'{mode} {variant} {reward}')
            assert done == k['done'], (mode, variant, reward, k['q'], k['v'])
            assert rew == pytest.approx(k['reward_value'], abs=1e-15)
            # host mirror of the Task
            hobs = task.observation_from_raw(k['q'], k['v'], k['a1'])
            np.testing.assert_allclose(hobs, k['obs'], rtol=0, atol=5e-16)
            hrew, hdone = task.get_state_info(hobs, [np.array(k['a0']), np.array(k['a1'])])
            assert hdone == k['done']
            assert hrew == pytest.approx(k['reward_value'], abs=1e-15)
            # device rule: raw thresholds
            dev_done = any((raw[c] < cfg.done_low[c]) or (raw[c] > cfg.done_high[c]) or np.isnan(raw[c])
                           for c in range(cfg.obs_dim))
            assert dev_done == k['done'], (mode, variant, reward, raw, k['q'], k['v'])
            n_done += k['done']
    assert n_done > 300


def test_leg_joint_angles(golden):
    task, compiled, cfg = make_config('fixed_hip')
    d = task.cfg.get_config('task_modes/fixed_hip/definition')
    for k in golden['leg_joint_angles']:
        hip, knee = oracle.leg_joint_angles(cfg, k['pitch'])
        assert hip == pytest.approx(k['hip'], abs=1e-14) and knee == pytest.approx(k['knee'], abs=1e-14)
        dd = dict(d)
        dd['planarizer_pitch_joint'] = k['pitch']
        h2, k2 = leg_joint_angles(dd)
        assert h2 == pytest.approx(k['hip'], abs=1e-14) and k2 == pytest.approx(k['knee'], abs=1e-14)
    # SURVEY.md section 8c known answers
    assert oracle.leg_joint_angles(cfg, 0.15) == pytest.approx([0.286105972506, -0.587730986633], abs=1e-11)
    assert oracle.leg_joint_angles(cfg, 0.2) == pytest.approx([0, 0])


def test_tolerance(golden):
    for k in golden['tolerance']:
        args = dict(bounds=tuple(k['bounds']), margin=k['margin'], sigmoid=k['sigmoid'], value_at_margin=k['value_at_margin'])
        assert oracle.tolerance(k['x'], **args) == pytest.approx(k['value'], abs=2e-15)
        assert rewards_utils.tolerance(k['x'], **args) == pytest.approx(k['value'], abs=2e-15)
    # vectorised + torch paths agree with the scalar path
    import torch
    xs = np.linspace(-1, 1, 21)
    a = rewards_utils.tolerance(xs, bounds=(0.25, 0.3), margin=0.15, sigmoid='tanh_squared')
    b = rewards_utils.tolerance(torch.as_tensor(xs), bounds=(0.25, 0.3), margin=0.15, sigmoid='tanh_squared')
    np.testing.assert_allclose(a, b.numpy(), atol=1e-15)
    with pytest.raises(ValueError):
        rewards_utils.tolerance(0.0, bounds=(1, 0))
    with pytest.raises(ValueError):
        rewards_utils.tolerance(0.0, margin=-1)
    with pytest.raises(ValueError):
        rewards_utils.tolerance(2.0, margin=1, sigmoid='nope')


def test_reward_support_lists():
    r = rewards.StraightV1({}, True)
    assert r.get_supported_task_modes() == ['simple'] and not r.is_task_supported('fixed_hip')
    assert rewards.BalancingV1({}, True).is_task_supported('free_hip')
    assert not rewards.HoppingV1({}, True).is_task_supported('simple')


def test_oracle_physics_regression_fixture():
    """The oracle's physics is pinned against ITSELF (tools/gen_oracle_regression.py): a regression guard for the
    restatement — touchdown, sliding contact, joint friction, implicit damping, both sweep-exit modes — not a parity
    claim (the reference holds no trajectory to pin it to)."""
    import json
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    import gen_oracle_regression as gen
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'oracle_physics_regression.json')
    data = json.load(open(path))
    touched = 0
    for case in data['cases']:
        got = gen.trajectory(case['mode'], case['reward'], case['reset'], case['pgs_tol'])
        for step, snap in case['snapshots'].items():
            np.testing.assert_allclose(got[step]['state'], snap['state'], rtol=0, atol=1e-10, err_msg=f"{case['mode']} {case['reset']} tol={case['pgs_tol']} step {step}")
            np.testing.assert_allclose(got[step]['obs'], snap['obs'], rtol=0, atol=1e-10)
            assert got[step]['reward'] == snap['reward']
        n = {'simple': 2, 'fixed': 3, 'fixed_hip': 4, 'free_hip': 5}[case['mode']]
        for snap in case['snapshots'].values():
            touched += int((np.array(snap["state"])[3 * n:-2:3] > 0).any())
    assert touched >= 4      # several snapshots catch a proxy pressing on the ground (the others are between bounces)
