"""CPU suite: host model compiler, task configuration, registry, randomizer wrappers' configuration,
vec-env helpers — everything on the host side of the C-ABI."""
import math

import numpy as np
import pytest

import gym_os2r_b200
from gym_os2r_b200 import _capi, _gymshim, rewards
from gym_os2r_b200.common.distributed import shard_range
from gym_os2r_b200.common.vec_env.cuda_vec_env import LazyInfos
from gym_os2r_b200.models import compiler, config
from gym_os2r_b200.tasks import monopod

from helpers import make_config


def _compile(name):
    return compiler.compile_model(name, config.SettingsConfig().get_config('physics'))


def test_chain_extraction_and_fixed_joint_lumping():
    full, fh, fx, sp = (_compile(n) for n in ('monopod', 'monopod-fixed_hip', 'monopod-fixed', 'monopod-simple'))
    assert full.joint_names == ['planarizer_yaw_joint', 'planarizer_pitch_joint', 'boom_connector_joint', 'hip_joint', 'knee_joint']
    assert fh.joint_names == ['planarizer_yaw_joint', 'planarizer_pitch_joint', 'hip_joint', 'knee_joint']
    assert fx.joint_names == ['planarizer_pitch_joint', 'hip_joint', 'knee_joint']
    assert sp.joint_names == ['hip_joint', 'knee_joint']
    # boom + hip_link lumped: one mass draw, COM between them (sdformat fixed-joint reduction)
    assert fh.body_names[1] == 'boom_link+hip_link'
    assert fh.struct.mass[1] == pytest.approx(0.692 + 0.151)
    assert list(full.struct.axis[:5]) == [2, 0, 0, 0, 0]
    assert fh.contact_names == ['hip', 'knee', 'foot'] and list(fh.struct.contact_body[:3]) == [2, 2, 3]
    # URDF dynamics: base model 0.01/0.001, fixed_hip 0/0.01
    assert full.struct.damping[0] == 0.01 and full.struct.friction[0] == 0.001
    assert fh.struct.damping[0] == 0.0 and fh.struct.friction[3] == 0.01


def test_lumped_inertia_is_parallel_axis_sum():
    fh, full = _compile('monopod-fixed_hip').struct, _compile('monopod').struct
    # total mass and first moment about the pitch joint are preserved by lumping at q_boom_connector = 0
    m_b, m_h = full.mass[1], full.mass[2]
    com_b = np.array(full.com[1][:])
    R = np.array(full.tree_R[2][:]).reshape(3, 3)
    com_h = np.array(full.tree_p[2][:]) + R @ np.array(full.com[2][:])
    expect = (m_b * com_b + m_h * com_h) / (m_b + m_h)
    np.testing.assert_allclose(np.array(fh.com[1][:]), expect, atol=1e-15)
    I = np.array(fh.inertia[1][:])
    assert I[0] > 0 and I[1] > 0 and I[2] > 0 and I[0] > 0.138   # dominated by the 2 m boom


def test_forward_kinematics_facts():
    """hip height = 0.11 + 2.01 sin(pitch) (+ offset along the skewed frame);
    nominal reset clearances of the foot: stand 0.048, half_stand 0.047, ground 0.003, float 0.137."""
    cm = _compile('monopod-fixed_hip')
    m = cm.struct
    for pitch, hip, knee, clearance in ((0.15, 0.286105972506, -0.587730986633, 0.048), (0.08, 0.923989186806, -1.92129571506, 0.047),
                                        (-0.005, 1.23412311, -2.69115382351, 0.003), (0.2, 0.0, 0.0, 0.137)):
        q = np.zeros(4)
        q[cm.dof_of('planarizer_pitch_joint')], q[cm.dof_of('hip_joint')], q[cm.dof_of('knee_joint')] = pitch, hip, knee
        Rs, ps, cs = compiler.forward_kinematics(m, q)
        assert ps[2][2] == pytest.approx(0.11 + 2.01 * math.sin(pitch), abs=0.008)
        assert cs[2][2] - m.contact_radius[2] == pytest.approx(clearance, abs=1.5e-3)
    # the literal rpy 1.57 (not pi/2) must survive: hip axis is skewed by ~8e-4 from the boom axis
    Rs, ps, cs = compiler.forward_kinematics(m, np.zeros(4))
    assert abs(Rs[2][0, 0]) == pytest.approx(7.96e-4, rel=0.02)


def test_contact_proxy_set_covers_the_links_that_can_reach_the_ground():
    """the reference collides every link's mesh (monopod.urdf:39-45,85-90,128-133,185-195,
    248-256). Which links can touch the ground plane at all is a geometric fact checked here from the mesh extents:
      * boom tube (r 0.0127 about the boom axis, y in [-2, 0]) and, where boom_connector is fixed, the hip bracket
        (bbox corner farthest from the boom axis: local (0.01, -0.015, 0.085)) are shielded by the hip sphere (r 0.02
        about the hip joint, 0.035 m outboard of the bracket): whatever the leg does, the sphere's lowest point is
        lower -> no proxy needed, none shipped;
      * with boom_connector FREE (`monopod`, task mode free_hip) the bracket turns about its own x axis and its far corner
        (0.107 m from that axis) swings far below the hip sphere -> the model carries a fourth proxy `hip_link`;
      * the central pivot's bottom face lies exactly in the ground plane and never translates: a zero-depth contact that
        carries no load (the pivot hangs from the world joint) -> ignored."""
    import itertools
    full, fh = _compile('monopod'), _compile('monopod-fixed_hip')
    assert full.contact_names == ['hip_link', 'hip', 'knee', 'foot'] and full.struct.n_contacts == 4
    assert list(full.struct.contact_body[:4]) == [2, 3, 3, 4]
    assert fh.contact_names == ['hip', 'knee', 'foot']
    # hip bracket bbox corners and boom tip in the frame of the body that carries them in the fixed_hip model (boom)
    m = full.struct
    R_bc = np.array(m.tree_R[2][:]).reshape(3, 3)
    p_bc = np.array(m.tree_p[2][:])
    corners = [p_bc + R_bc @ np.array(c) for c in itertools.product((-0.01, 0.01), (-0.015, 0.065), (-0.015, 0.085))]
    tip = np.array([0.0, -2.0, 0.0])
    rng = np.random.RandomState(0)
    for _ in range(300):
        # boom pitched low enough for anything at its end to come within 5 cm of the ground
        q = np.array([rng.uniform(-3, 3), rng.uniform(-0.07, -0.02), rng.uniform(-3.2, 3.2), rng.uniform(-3.2, 3.2)])
        Rs, ps, cs = compiler.forward_kinematics(fh.struct, q)
        hip_bottom = cs[0][2] - fh.struct.contact_radius[0]
        lowest_bracket = min((ps[1] + Rs[1] @ c)[2] for c in corners)
        boom_tip_bottom = (ps[1] + Rs[1] @ tip)[2] - 0.0127
        assert lowest_bracket - hip_bottom > 0.004 and boom_tip_bottom - hip_bottom > 0.006, (q, lowest_bracket, boom_tip_bottom, hip_bottom)
    # free boom_connector: turned by about +-2 rad the bracket corner is the lowest point of the whole robot
    worst = 0.0
    for bc in np.linspace(-3.1, 3.1, 63):
        q = np.array([0.0, 0.0, bc, 0.5, -1.0])
        Rs, ps, cs = compiler.forward_kinematics(m, q)
        worst = max(worst, (cs[1][2] - m.contact_radius[1]) - (cs[0][2] - m.contact_radius[0]))
    assert worst > 0.06          # the bracket proxy reaches > 6 cm below the hip sphere


def test_unsupported_urdf_features_raise(tmp_path):
    text = open(gym_os2r_b200.models.assets.get_model_file('monopod')).read()
    bad = tmp_path / 'bad.urdf'
    bad.write_text(text.replace('<axis xyz="0 0 1"/>', '<axis xyz="0 0.6 0.8"/>'))
    with pytest.raises(ValueError):
        compiler.compile_urdf(str(bad), config.SettingsConfig().get_config('physics'))
    with pytest.raises(RuntimeError):
        gym_os2r_b200.models.assets.get_model_file('old-monopod')


def test_settings_config_get_set_isolated():
    cfg = config.SettingsConfig()
    a = cfg.get_config('task_modes/fixed_hip/spaces/observation/hip_joint/limits')
    a[0] = 123.0                                   # deep copy: live config untouched
    assert cfg.get_config('task_modes/fixed_hip/spaces/observation/hip_joint/limits')[0] == 6.28319
    cfg.set_config(0.3, 'resets/stand/planarizer_pitch_joint')
    assert cfg.get_config('/resets/stand')['planarizer_pitch_joint'] == 0.3
    # anchors were expanded into independent containers
    cfg.set_config([1, -1, 1, -1], 'task_modes/fixed/spaces/observation/hip_joint/limits')
    assert cfg.get_config('task_modes/fixed_hip/spaces/observation/hip_joint/limits')[0] == 6.28319
    with pytest.raises(KeyError):
        cfg.get_config('task_modes/old-free_hip')


def test_task_validation_errors():
    T = monopod.MonopodTask
    with pytest.raises(RuntimeError, match='Missing required kwarg'):
        T(1000, task_mode='fixed_hip', reward_class=rewards.BalancingV1)
    with pytest.raises(RuntimeError, match='reset positions'):
        T(1000, task_mode='fixed_hip', reward_class=rewards.BalancingV1, reset_positions=['nope'])
    with pytest.raises(RuntimeError, match='not supported'):
        T(1000, task_mode='old-fixed', reward_class=rewards.BalancingV1, reset_positions=['stand'])
    t = T(1000, task_mode='fixed_hip', reward_class=rewards.StraightV1, reset_positions=['stand'])
    with pytest.raises(AssertionError):
        t.create_spaces()
    with pytest.warns(SyntaxWarning):
        T(1000, task_mode='fixed_hip', reward_class=rewards.BalancingV1, reset_positions=['stand'], extra=1)


def test_custom_settings_flow_into_device_config():
    """examples/minimal_example_testing.py:17-31 style: user-edited reset pose via SettingsConfig."""
    cfg = config.SettingsConfig()
    cfg.set_config({'laying_down': False, 'planarizer_pitch_joint': 0.3}, 'resets/high')
    task, cm, tc = make_config('fixed_hip', reset_positions=('high', 'lay'), config=cfg)
    assert tc.n_resets == 2 and tc.reset_pitch[0] == 0.3 and tc.reset_laying[1] == 1
    assert tc.ik_boom == 2100 and tc.ik_clip == 25


def test_done_thresholds_agree_with_reference_formula():
    """Raw-unit thresholds == `not reset_space.contains(normalised obs)` for values hugging the limits."""
    task, cm, tc = make_config('free_hip', reward='BalancingV1')
    rng = np.random.RandomState(0)
    for col in range(tc.obs_dim):
        kind = tc.obs_kind[col]
        if kind == _capi.OBS_POS_PERIODIC:
            continue
        centre = tc.done_high[col]
        for x in np.concatenate([centre + rng.uniform(-1e-9, 1e-9, 50), [centre, np.nextafter(centre, np.inf), np.nextafter(centre, -np.inf)],
                                 -centre + rng.uniform(-1e-9, 1e-9, 50)]):
            norm = math.tanh(0.05 * x) if kind == _capi.OBS_VEL else 2 * (x - tc.obs_low[col]) / (tc.obs_high[col] - tc.obs_low[col]) - 1
            ref_done = not (task.reset_space.low[col] <= norm <= task.reset_space.high[col])
            dev_done = x < tc.done_low[col] or x > tc.done_high[col]
            assert ref_done == dev_done, (col, x)


def test_registered_ids_and_kwargs():
    """gym_os2r/__init__.py:16-128 — ids, task modes, reward classes, reset positions, time limits."""
    specs = {s.id: s for s in _gymshim.registry.all()}
    expect = {'Monopod-stand-v1': ('fixed_hip', 'StandingV1', ['ground'], 100000),
              'Monopod-balance-v1': ('fixed_hip_simple', 'BalancingV1', ['stand'], 100000),
              'Monopod-balance-v2': ('fixed_hip_simple', 'BalancingV2', ['stand'], 100000),
              'Monopod-balance-v3': ('fixed_hip_simple', 'BalancingV2', ['stand', 'half_stand', 'ground', 'lay', 'float'], 10000),
              'Monopod-nonorm-balance-v1': ('fixed_hip_simple', 'BalancingV1', ['stand'], 100000),
              'Monopod-nonorm-balance-v2': ('fixed_hip_simple', 'BalancingV2', ['stand'], 100000),
              'Monopod-nonorm-balance-v3': ('fixed_hip_simple', 'BalancingV2', ['stand', 'half_stand', 'ground', 'lay', 'float'], 10000),
              'Monopod-hop-v1': ('free_hip', 'HoppingV1', ['stand'], 100000),
              'Monopod-simple-v1': ('simple', 'StraightV1', ['stand'], 100000)}
    assert set(specs) == set(expect)
    for eid, (mode, rew, poses, limit) in expect.items():
        s = specs[eid]
        assert s.kwargs['task_mode'] == mode and s.kwargs['reward_class'].__name__ == rew
        assert s.kwargs['reset_positions'] == poses and s.max_episode_steps == limit
        assert s.kwargs['agent_rate'] == 1000 and s.kwargs['physics_rate'] == 10000
        assert ('no_norm' in s.kwargs['task_cls'].__module__) == ('nonorm' in eid)


def test_runtime_construction_and_wrapper_configuration():
    """Construction needs no GPU (the engine is created lazily): spaces, obs dims (tests_general.py:87,111),
    and what the two randomizer wrappers configure."""
    from gym_os2r_b200 import randomizers
    from gym_os2r_b200.common import make_env_from_id
    import functools
    dims = {'Monopod-stand-v1': 8, 'Monopod-balance-v1': 5, 'Monopod-hop-v1': 10, 'Monopod-simple-v1': 4}
    for eid, d in dims.items():
        env = randomizers.monopod.MonopodEnvRandomizer(env=functools.partial(make_env_from_id, env_id=eid))
        rt = env.unwrapped
        assert env.observation_space.shape == (d,) and env.action_space.shape == (2,)
        assert env.observation_space.dtype == np.float64
        assert rt._cfg.reset_randomized == 1 and rt._cfg.randomize_params == 1 and rt._cfg.randomize_gravity == 1
        assert rt._cfg.mu_link == 0.33 and rt._cfg.fric_hi == 0.05
        assert rt.num_of_steps_per_run == 10 and rt._compiled.struct.substeps == 10
        assert rt._cfg.max_episode_steps == 100000 and rt._cfg.auto_reset == 0
    env = randomizers.monopod_no_rand.MonopodEnvNoRandomizer(
        env=functools.partial(make_env_from_id, env_id='Monopod-balance-v1', task_mode='free_hip'))
    assert env.observation_space.shape == (10,) and env.unwrapped._cfg.reset_randomized == 0
    env = randomizers.monopod_no_rand.MonopodEnvNoRandomizer(
        env=functools.partial(make_env_from_id, env_id='Monopod-simple-v1'))
    assert env.unwrapped._cfg.simple_sample_reset == 1 and env.unwrapped._cfg.simple_hi[0] == 1.0
    assert env.get_state_info(np.zeros(4), [0, 0])[1] is False


def test_get_state_info_accepts_bare_action_and_history():
    task, cm, tc = make_config('fixed_hip', reward='BalancingV3')
    obs = np.zeros(8)
    obs[task.observation_index['planarizer_pitch_joint_pos']] = 0.1
    r1, d1 = task.get_state_info(obs, [0.2, -0.1])
    r2, d2 = task.get_state_info(obs, [np.array([0.2, -0.1]), np.array([0.2, -0.1])])
    assert r1 == r2 and d1 is False and d2 is False
    batch = np.tile(obs, (5, 1))
    batch[2, 0] = 1.5
    rb, db = task.get_state_info(batch, np.zeros((5, 2)))
    assert db.tolist() == [False, False, True, False, False] and rb.shape == (5,)


def test_custom_reward_subclass_is_flagged_for_torch_evaluation():
    class ExampleV0(rewards.RewardBase):          # examples/minimal_example.py:15-29
        def __init__(self, observation_index, normalized):
            super().__init__(observation_index, normalized)
            self.supported_task_modes = ['fixed_hip']

        def calculate_reward(self, obs, actions):
            return 1
    import warnings
    from gym_os2r_b200.runtimes.configure import configure
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        task, cm, tc = configure(monopod.MonopodTask, task_mode='fixed_hip', reward_class=ExampleV0, reset_positions=['stand'])
    assert tc.reward_id == _capi.REWARD_CUSTOM


def test_shard_range_and_lazy_infos():
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert shard_range(1 << 20, 7, 8) == (7 * 131072, 131072)
    dense = LazyInfos.from_dense(['stand', 'lay'], np.array([0, 1, 1]), np.array([False, True, True]),
                                 np.arange(6.0).reshape(3, 2), np.array([0, 1, 2]))
    # the sparse form the packed host step produces: one record per finished env, in any order
    sparse = LazyInfos(['stand', 'lay'], np.array([0, 1, 1], dtype=np.uint8), np.array([False, True, True]),
                       np.array([2, 1]), np.array([2, 1]), np.array([[4.0, 5.0], [2.0, 3.0]]))
    for infos in (dense, sparse):
        assert len(infos) == 3 and infos[0] == {'reset_orientation': 'stand'}
        assert infos[2]['TimeLimit.truncated'] and infos[1]['terminal_observation'].tolist() == [2.0, 3.0]
        assert 'TimeLimit.truncated' not in infos[1] and infos[2]['terminal_observation'].tolist() == [4.0, 5.0]
        assert [i['reset_orientation'] for i in infos] == ['stand', 'lay', 'lay']
        assert sorted(infos.terminal_indices.tolist()) == [1, 2]


def test_numa_binding_helper_never_fails_and_never_widens_the_mask():
    """bind_to_gpu_numa_node: without NVML / a GPU it reports why and leaves the affinity alone."""
    import os
    from gym_os2r_b200.common.distributed import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    out = bind_to_gpu_numa_node(0)
    assert isinstance(out, dict) and 'bound' in out
    assert os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)


def test_randomizer_wrapper_forwards_num_physics_rollouts():
    """MonopodEnvRandomizer(num_physics_rollouts=K) -> os2r_task_cfg.gravity_redraw_resets (randomizers/monopod.py:36,371)."""
    import functools
    from gym_os2r_b200 import randomizers
    from gym_os2r_b200.common import make_env_from_id
    env = randomizers.monopod.MonopodEnvRandomizer(env=functools.partial(make_env_from_id, env_id='Monopod-balance-v1'),
                                                   num_physics_rollouts=4)
    cfg = env.unwrapped._cfg
    assert cfg.gravity_redraw_resets == 4 and cfg.randomize_gravity == 1 and cfg.reset_randomized == 1
    with pytest.raises(ValueError):
        randomizers.monopod.MonopodEnvRandomizer(env=functools.partial(make_env_from_id, env_id='Monopod-balance-v1'),
                                                 num_physics_rollouts=-1)


def test_solver_constants_reach_the_model_struct():
    """The backend's own solver constants (settings.yaml `physics:`) are forwarded into os2r_model, and the per-runtime
    overrides of configure() win: sweep cap, per-env exit tolerance, sweeps that include the joint-friction rows."""
    from helpers import make_config
    task, cm, cfg = make_config('fixed_hip', pgs_tol=None)
    assert cm.struct.pgs_iters == 8 and cm.struct.pgs_tol == 1e-6 and cm.struct.pgs_joint_sweeps == 1
    task, cm, cfg = make_config('fixed_hip', pgs_iters=5, pgs_tol=0.0, pgs_joint_sweeps=0)
    assert cm.struct.pgs_iters == 5 and cm.struct.pgs_tol == 0.0 and cm.struct.pgs_joint_sweeps == 0
