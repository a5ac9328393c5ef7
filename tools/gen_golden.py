#!/usr/bin/env python
"""Generate tests/golden/task_kat.json from the REFERENCE's own numpy code.

Runs only in the build container (needs /root/reference, read-only); the GPU box and the test
suite consume the committed JSON. The reference modules are loaded *unmodified* from
/root/reference with ~30 lines of stub modules standing in for gym / gym_ignition / scenario
(SURVEY.md appendix A): tasks/monopod.py, tasks/monopod_no_norm.py, rewards/__init__.py,
rewards/rewards_utils.py, utils/reset.py, models/config/__init__.py (+ settings.yaml).

What is recorded (all float64, exact repr):
  * spaces: per task mode x {norm, no_norm}: observation_index, observation_mask, periodic
    columns, obs limits, obs/reset space bounds;
  * task KATs: (joint positions, velocities in the task's joint_names order, a_t, a_{t-1}) ->
    observation, reward, done through task.set_action / get_observation / get_reward / is_done,
    for every supported reward class; random states plus termination / wrap edge cases;
  * leg_joint_angles(pitch) sweep; tolerance() for all 8 sigmoids.
"""
import importlib.util
import json
import os
import sys
import types
from collections import deque

import numpy as np

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_os2r_b200._gymshim import Box  # noqa: E402  (a plain Box with gym's contains())


def _stub_modules():
    gym = types.ModuleType('gym')
    gym.spaces = types.ModuleType('gym.spaces')
    gym.spaces.Box = Box
    gi = types.ModuleType('gym_ignition')
    base = types.ModuleType('gym_ignition.base')
    task_mod = types.ModuleType('gym_ignition.base.task')

    class Task:
        def __init__(self, agent_rate):
            self.agent_rate = agent_rate
    task_mod.Task = Task
    base.task = task_mod
    utils = types.ModuleType('gym_ignition.utils')
    logger = types.ModuleType('gym_ignition.utils.logger')
    logger.debug = logger.warn = logger.info = lambda *a, **k: None
    utils.logger = logger
    typing_mod = types.ModuleType('gym_ignition.utils.typing')
    for name in ('Action', 'Reward', 'ActionSpace', 'ObservationSpace'):
        setattr(typing_mod, name, object)
    typing_mod.Observation = np.array
    utils.typing = typing_mod
    gi.base, gi.utils = base, utils
    scen = types.ModuleType('scenario')
    core = types.ModuleType('scenario.core')
    core.JointControlMode_force = 1
    scen.core = core
    mods = {'gym': gym, 'gym.spaces': gym.spaces, 'gym_ignition': gi, 'gym_ignition.base': base,
            'gym_ignition.base.task': task_mod, 'gym_ignition.utils': utils,
            'gym_ignition.utils.logger': logger, 'gym_ignition.utils.typing': typing_mod,
            'scenario': scen, 'scenario.core': core}
    sys.modules.update(mods)
    for pkg in ('gym_os2r', 'gym_os2r.models', 'gym_os2r.rewards', 'gym_os2r.tasks', 'gym_os2r.utils'):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, *pkg.split('.'))]
        sys.modules[pkg] = m


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class FakeModel:
    """Stands in for the ScenarIO model: serves joint state, echoes force targets."""

    def __init__(self):
        self.pos, self.vel, self.targets = {}, {}, {}

    def joint_positions(self, names):
        return [self.pos[n] for n in names]

    def joint_velocities(self, names):
        return [self.vel[n] for n in names]

    def set_joint_generalized_force_targets(self, data, names):
        self.targets = dict(zip(names, [float(x) for x in data]))
        return True

    def joint_generalized_force_targets(self, names):
        return [self.targets[n] for n in names]


def main():
    _stub_modules()
    _load('gym_os2r.models.config', 'gym_os2r/models/config/__init__.py')
    ru = _load('gym_os2r.rewards.rewards_utils', 'gym_os2r/rewards/rewards_utils.py')
    rw = _load('gym_os2r.rewards', 'gym_os2r/rewards/__init__.py')
    t_norm = _load('gym_os2r.tasks.monopod', 'gym_os2r/tasks/monopod.py')
    t_raw = _load('gym_os2r.tasks.monopod_no_norm', 'gym_os2r/tasks/monopod_no_norm.py')
    rs = _load('gym_os2r.utils.reset', 'gym_os2r/utils/reset.py')

    rng = np.random.RandomState(20261018)
    out = {'generator': 'tools/gen_golden.py', 'reference': 'OpenSim2Real/gym-os2r 1.2.0 (/root/reference)',
           'spaces': [], 'task_kat': [], 'leg_joint_angles': [], 'tolerance': []}
    modes = ['free_hip', 'fixed_hip', 'fixed_hip_torque', 'fixed_hip_simple', 'fixed', 'simple']
    reward_names = ['BalancingV1', 'BalancingV2', 'BalancingV3', 'StandingV1', 'HoppingV1', 'StraightV1']
    pos_limit = {'hip_joint': 6.28319, 'knee_joint': np.pi, 'planarizer_pitch_joint': 1.5708,
                 'planarizer_yaw_joint': np.pi, 'boom_connector_joint': 6.28319}

    for mode in modes:
        for variant, mod in (('norm', t_norm), ('no_norm', t_raw)):
            for rname in reward_names:
                rcls = getattr(rw, rname)
                probe = rcls({}, True)
                if mode not in probe.get_supported_task_modes():
                    continue
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    task = mod.MonopodTask(agent_rate=1000, task_mode=mode, reward_class=rcls,
                                           reset_positions=['stand'])
                if rname == 'HoppingV1' and 'planarizer_yaw_joint_vel' in task.observation_name_mask:
                    continue   # reference would KeyError: reward needs a masked column
                if rname == 'HoppingV1' and 'planarizer_yaw_joint' not in task.joint_names:
                    continue
                aspace, ospace = task.create_spaces()
                task.action_space, task.observation_space = aspace, ospace
                task.model = FakeModel()
                if rname in ('BalancingV1', 'StraightV1'):
                    mask_attr = 'observation_mask' if variant == 'norm' else 'observaton_mask'
                    out['spaces'].append({
                        'task_mode': mode, 'variant': variant,
                        'joint_names': task.joint_names, 'action_names': task.action_names,
                        'observation_index': task.observation_index,
                        'observation_mask': [int(i) for i in getattr(task, mask_attr)],
                        'periodic_joints': [int(i) for i in task.periodic_joints],
                        'obs_low': ospace.low.tolist(), 'obs_high': ospace.high.tolist(),
                        'reset_low': task.reset_space.low.tolist(), 'reset_high': task.reset_space.high.tolist(),
                        'max_torques': task.max_torques.tolist()})
                names = task.joint_names
                cases = []
                for _ in range(14):     # generic random states
                    q = [float(rng.uniform(-0.95, 0.95) * pos_limit[n]) for n in names]
                    v = [float(rng.normal(0, 8.0)) for n in names]
                    cases.append((q, v))
                for _ in range(4):      # pitch inside / around the reward band, small velocities
                    q = [float(rng.uniform(-0.5, 0.5)) for n in names]
                    if 'planarizer_pitch_joint' in names:
                        q[names.index('planarizer_pitch_joint')] = float(rng.uniform(0.05, 0.5))
                    v = [float(rng.normal(0, 0.5)) for n in names]
                    cases.append((q, v))
                for _ in range(3):      # wrapped periodic joints, several turns
                    q = [float(rng.uniform(-0.9, 0.9) * pos_limit[n]) for n in names]
                    for j, n in enumerate(names):
                        if n in ('knee_joint', 'planarizer_yaw_joint'):
                            q[j] = float(rng.uniform(-40, 40))
                    v = [float(rng.normal(0, 3.0)) for n in names]
                    cases.append((q, v))
                # termination edges: each non-periodic position just inside / outside its limit,
                # each velocity around the tanh saturation point, knee exactly at +-pi
                for j, n in enumerate(names):
                    if n in ('knee_joint', 'planarizer_yaw_joint'):
                        for val in (np.pi, -np.pi, 3 * np.pi):
                            q = [0.1] * len(names); q[j] = float(val)
                            cases.append((q, [0.0] * len(names)))
                        continue
                    for sgn in (1, -1):
                        for delta in (-1e-9, 0.0, 1e-9):
                            q = [0.1] * len(names); q[j] = sgn * (pos_limit[n] + delta)
                            cases.append((q, [0.0] * len(names)))
                for j, n in enumerate(names):
                    for val in (360.0, -360.0, 375.0, -375.0, 172.0, 174.0):
                        v = [0.0] * len(names); v[j] = val
                        cases.append(([0.1] * len(names), v))
                for q, v in cases:
                    a1 = rng.uniform(-1, 1, 2)
                    a0 = rng.uniform(-1, 1, 2)
                    if rng.rand() < 0.25:
                        a0 = a1 + rng.uniform(-0.08, 0.08, 2)   # exercise the margin-0.1 band of HoppingV1
                        a0 = np.clip(a0, -1, 1)
                    task.model.pos = dict(zip(names, q))
                    task.model.vel = dict(zip(names, v))
                    task.action_history = deque([a1.copy()] + [np.zeros(2) for _ in range(9)], maxlen=10)
                    task.set_action(a0, store_action=True)
                    obs = task.get_observation()
                    rew = task.get_reward()
                    done = task.is_done()
                    out['task_kat'].append({
                        'task_mode': mode, 'variant': variant, 'reward': rname,
                        'q': q, 'v': v, 'a0': a0.tolist(), 'a1': a1.tolist(),
                        'obs': [float(x) for x in obs], 'reward_value': float(rew), 'done': bool(done)})

    definition = {'upper_leg_length': 200, 'lower_leg_length': 190, 'central_pivot_height': 80,
                  'length_boom': 2100, 'hip_offset': 0, 'clipping_adjust': 25}
    for bp in list(np.linspace(-0.0065, 0.26, 60)) + [0.15, 0.08, -0.005, 0.2, 0.12, 0.18]:
        d = dict(definition)
        d['planarizer_pitch_joint'] = float(bp)
        ang = rs.leg_joint_angles(d)
        out['leg_joint_angles'].append({'pitch': float(bp), 'hip': float(ang[0]), 'knee': float(ang[1])})

    sig_cases = [('gaussian', 0.1), ('hyperbolic', 0.25), ('long_tail', 0.1), ('reciprocal', 0.1), ('cosine', 0.0),
                 ('cosine', 0.3), ('linear', 0.1), ('linear', 0.0), ('quadratic', 0.4), ('quadratic', 0.0),
                 ('tanh_squared', 0.1)]
    for sig, vam in sig_cases:
        for bounds, margin in (((0.0, 0.0), 1.0), ((0.25, 0.3), 0.15), ((0.07, 0.28), 0.01), ((-0.5, 0.5), 0.0)):
            for x in np.concatenate((np.linspace(-1.5, 1.5, 13), [0.25, 0.3, 0.07, 0.28, 0.06])):
                val = ru.tolerance(float(x), bounds=bounds, margin=margin, sigmoid=sig, value_at_margin=vam)
                out['tolerance'].append({'x': float(x), 'bounds': list(bounds), 'margin': margin,
                                         'sigmoid': sig, 'value_at_margin': vam, 'value': float(val)})

    path = os.path.join(ROOT, 'tests', 'golden', 'task_kat.json')
    with open(path, 'w') as f:
        json.dump(out, f, separators=(',', ':'))
    print(f'wrote {path}: {len(out["task_kat"])} task KATs, {len(out["spaces"])} spaces, '
          f'{len(out["leg_joint_angles"])} IK, {len(out["tolerance"])} tolerance; '
          f'{os.path.getsize(path) / 1e6:.2f} MB')


if __name__ == '__main__':
    main()
