"""Reset poses pinned to the REFERENCE's own code (CPU suite).

tests/golden/reset_poses.npz holds joint positions recorded by running the unmodified
``MonopodRandomizersMixin.randomize_task`` (gym_os2r/randomizers/monopod.py:67-135) and
``MonopodEnvNoRandomizer.randomize_task`` (monopod_no_rand.py:26-98) under stubbed simulator objects
(tools/gen_reset_golden.py). The reference draws from numpy's global RNG, this backend from one Philox stream per env:
the two are distribution-equal, not stream-equal, so the randomised poses are compared by two-sample Kolmogorov-Smirnov
tests per joint plus the exact discrete structure (mirror direction, lay side, the sign-precedence quirk of :102-103,
the unperturbed knee); the NoRandomizer poses must match to rounding. The oracle is what is compared here; the GPU
suite checks device == oracle to 1e-12 and repeats the KS test on device draws."""
import os

import numpy as np
import pytest
from scipy import stats as sps

import oracle

from helpers import make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POSES = ['stand', 'half_stand', 'ground', 'lay', 'float']


@pytest.fixture(scope='module')
def poses():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'reset_poses.npz'))


def joint_order_positions(task, cm, state):
    """oracle / engine state rows (chain order) -> columns in task.joint_names order (the fixture's order)"""
    return np.stack([state[:, cm.dof_of(name)] for name in task.joint_names], 1)


def check_randomized_pose_distribution(ref, got, joint_names, pose, pmin=1e-3):
    """ref / got: [n, n_joints] in joint_names order."""
    got = got.astype(np.float32).astype(np.float64)      # the fixture is stored as float32: compare like with like (atoms!)
    col = {n: i for i, n in enumerate(joint_names)}
    hip_r, hip_g = ref[:, col['hip_joint']], got[:, col['hip_joint']]
    knee_r, knee_g = ref[:, col['knee_joint']], got[:, col['knee_joint']]
    for name in ('planarizer_pitch_joint', 'planarizer_yaw_joint'):
        assert sps.ks_2samp(ref[:, col[name]], got[:, col[name]]).pvalue > pmin, (pose, name)
    if 'boom_connector_joint' in col:
        assert not ref[:, col['boom_connector_joint']].any() and not got[:, col['boom_connector_joint']].any()
    if pose == 'float':          # IK returns (0, 0) above the reachable height; (a>0) is false: no noise at all
        assert not hip_r.any() and not knee_r.any() and not hip_g.any() and not knee_g.any()
        return
    # mirror direction: P(dir = +1) = 1/2 ; hip sign = dir wherever the hip is non-zero (the IK returns (0, 0) — and
    # no noise is added — when the randomised pitch lifts the hip out of reach: part of the `stand` draws)
    assert abs((hip_r == 0).mean() - (hip_g == 0).mean()) < 0.02, pose
    for h in (hip_r, hip_g):
        assert abs((h[h != 0] > 0).mean() - 0.5) < 0.025, pose
    assert sps.ks_2samp(np.abs(hip_r), np.abs(hip_g)).pvalue > pmin, pose
    assert sps.ks_2samp(hip_r, hip_g).pvalue > pmin, pose
    if pose == 'lay':
        # laying: hip = +-1.57 (p = 1/2); only the POSITIVE side gets the noise (precedence quirk of :108), knee = 0
        assert not knee_r.any() and not knee_g.any()
        for h in (hip_r, hip_g):
            exact = np.abs(np.abs(h) - 1.57) < 1e-6
            assert abs(exact.mean() - 0.5) < 0.02
            assert (np.abs(h)[~exact] > 1.57).all()
    else:
        # knee <= 0 from the IK, so (knee>0 - knee<0) == (knee>0) == 0: the knee is NEVER perturbed; it is the
        # mirrored IK value of the randomised pitch. |knee| is then a monotone function of the pitch draw.
        assert np.array_equal(hip_r == 0, knee_r == 0) and np.array_equal(hip_g == 0, knee_g == 0)
        assert sps.ks_2samp(np.abs(knee_r), np.abs(knee_g)).pvalue > pmin, pose
        assert np.array_equal(np.sign(knee_r), -np.sign(hip_r)) and np.array_equal(np.sign(knee_g), -np.sign(hip_g))


@pytest.mark.parametrize('mode', ['fixed_hip', 'free_hip'])
def test_randomizer_reset_poses_match_the_reference_distribution(poses, mode):
    names = [str(n) for n in poses[f'rand/{mode}/joint_names']]
    for k, pose in enumerate(POSES):
        ref = poses[f'rand/{mode}/{pose}'].astype(np.float64)
        task, cm, cfg = make_config(mode, reward='BalancingV1', reset_positions=(pose,), reset_randomized=True)
        assert task.joint_names == names
        orc = oracle.Oracle(cm.struct, cfg, len(ref), seed=31 + k, first_env_id=12345)
        orc.reset()
        assert not orc.qd.any()
        check_randomized_pose_distribution(ref, joint_order_positions(task, cm, orc.state), names, pose)


def test_reset_position_choice_is_uniform_like_the_reference(poses):
    """np.random.choice(task.reset_positions) (:89): every position with equal probability; per-position poses of the
    mixed run follow the single-position distributions."""
    ref_id = poses['rand/fixed_hip/all/chosen'].astype(int)
    task, cm, cfg = make_config('fixed_hip', reward='BalancingV1', reset_positions=tuple(POSES), reset_randomized=True)
    orc = oracle.Oracle(cm.struct, cfg, len(ref_id), seed=5)
    orc.reset()
    counts_ref = np.bincount(ref_id, minlength=5)
    counts_got = np.bincount(orc.reset_id, minlength=5)
    assert sps.chisquare(counts_got).pvalue > 1e-3 and sps.chisquare(counts_ref).pvalue > 1e-3
    assert sps.chi2_contingency(np.stack([counts_ref, counts_got]))[1] > 1e-3
    names = [str(n) for n in poses['rand/fixed_hip/joint_names']]
    got = joint_order_positions(task, cm, orc.state)
    ref = poses['rand/fixed_hip/all'].astype(np.float64)
    for k, pose in enumerate(POSES):
        check_randomized_pose_distribution(ref[ref_id == k], got[orc.reset_id == k], names, pose, pmin=1e-4)


@pytest.mark.parametrize('mode', ['fixed_hip', 'free_hip', 'fixed', 'fixed_hip_simple'])
def test_norandomizer_poses_equal_the_reference(poses, mode):
    """monopod_no_rand.py:60-84: nominal pitch, IK leg angles (or (1.57, 0) laying), everything else 0 — exactly."""
    names = [str(n) for n in poses[f'norand/{mode}/joint_names']]
    for pose in POSES:
        ref = poses[f'norand/{mode}/{pose}']
        task, cm, cfg = make_config(mode, reward='BalancingV1', reset_positions=(pose,), reset_randomized=False)
        assert task.joint_names == names
        orc = oracle.Oracle(cm.struct, cfg, 4, seed=1)
        orc.reset()
        got = joint_order_positions(task, cm, orc.state)
        np.testing.assert_allclose(got, np.repeat(ref, 4, 0), rtol=0, atol=1e-14)


def test_norandomizer_simple_mode_samples_the_observation_space(poses):
    """monopod_no_rand.py:84: hip, knee = observation_space.sample() columns -> U(low, high) per joint."""
    ref = poses['norand/simple/stand'].astype(np.float64)
    task, cm, cfg = make_config('simple', reward='StraightV1', reset_positions=('stand',), reset_randomized=False)
    orc = oracle.Oracle(cm.struct, cfg, len(ref), seed=9)
    orc.reset()
    got = joint_order_positions(task, cm, orc.state)
    lo, hi = task.observation_space.low, task.observation_space.high
    for j, name in enumerate(('hip_joint_pos', 'knee_joint_pos')):
        c = task.observation_index[name]
        assert sps.ks_2samp(ref[:, j], got[:, j]).pvalue > 1e-3
        assert sps.kstest(got[:, j], sps.uniform(lo[c], hi[c] - lo[c]).cdf).pvalue > 1e-3
