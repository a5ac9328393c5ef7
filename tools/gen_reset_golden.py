#!/usr/bin/env python
"""Generate tests/golden/reset_poses.npz by running the REFERENCE's own reset code.

Runs only in the build container (needs /root/reference, read-only). The reference files are loaded *unmodified*:
``gym_os2r/randomizers/monopod.py`` (``MonopodRandomizersMixin.randomize_task``, :67-135) and
``gym_os2r/randomizers/monopod_no_rand.py`` (``MonopodEnvNoRandomizer.randomize_task``, :26-98), together with the
task / config / IK modules tools/gen_golden.py already loads. Everything they import from the absent native stack
(lxml, scenario.gazebo, gym_ignition.randomizers.*, gym_ignition.scenario.*) is replaced by stub modules, and the
world / model / simulator objects by fakes that RECORD the joint positions handed to ``reset_joint_positions``.
The only methods overridden on the mixin are ``randomize_model_description`` / ``randomize_ground_description``
(SDF string generation through gym-ignition's SDFRandomizer + lxml: un-vendored, not on the reset-pose path).

Recorded under a seeded global ``np.random`` (the reference draws reset poses from the GLOBAL numpy RNG, :89-112):
  * ``rand/<mode>/<pose>``: float32 [10000, n_joints] joint positions (task.joint_names order) per reset position,
    MonopodEnvRandomizer path; plus the mixed ``rand/<mode>/all`` run with all five positions and its chosen names;
  * ``norand/<mode>/<pose>``: the deterministic poses of the NoRandomizer path (float64), and for ``simple`` 10000
    samples of the observation-space draw (:84).
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import gen_golden as gg  # noqa: E402  (stub modules + loader shared with the task KAT generator)

REF = gg.REF
POSES = ['stand', 'half_stand', 'ground', 'lay', 'float']
N_SAMPLES = 10000


def _extra_stubs():
    """Stand-ins for what the two randomizer files import beyond tools/gen_golden.py's stubs."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    lxml = mod('lxml')
    lxml.etree = mod('lxml.etree')
    scen = sys.modules['scenario']
    scen.gazebo = mod('scenario.gazebo', PhysicsEngine_dart=1,
                      urdffile_to_sdfstring=lambda f: (_ for _ in ()).throw(RuntimeError('not on the reset-pose path')))
    gi = sys.modules['gym_ignition']

    class _Base:
        def __init__(self, *a, **k):
            pass

    class PhysicsRandomizer(_Base):
        def __init__(self, randomize_after_rollouts_num=0):
            self.randomize_after_rollouts_num = randomize_after_rollouts_num

    class TaskRandomizer(_Base):
        pass

    class ModelDescriptionRandomizer(_Base):
        pass
    abc_mod = mod('gym_ignition.randomizers.abc', TaskRandomizer=TaskRandomizer, PhysicsRandomizer=PhysicsRandomizer,
                  ModelDescriptionRandomizer=ModelDescriptionRandomizer)

    class GazeboEnvRandomizer(_Base):
        pass
    ger = mod('gym_ignition.randomizers.gazebo_env_randomizer', GazeboEnvRandomizer=GazeboEnvRandomizer,
              MakeEnvCallable=object)
    sdf = mod('gym_ignition.randomizers.model.sdf', Method=object, Distribution=object, UniformParams=object,
              SDFRandomizer=object)
    model = mod('gym_ignition.randomizers.model', sdf=sdf)
    gi.randomizers = mod('gym_ignition.randomizers', abc=abc_mod, gazebo_env_randomizer=ger, model=model)
    gi.utils.misc = mod('gym_ignition.utils.misc', string_to_file=lambda s: s)

    class FakeMonopod:
        """gym_os2r/models/monopod.py:10-38 without the simulator: registers itself in the fake world."""

        def __init__(self, world, monopod_version, position=(0, 0, 0), orientation=(1, 0, 0, 0), model_file=None):
            self._name = 'monopod'
            self.version = monopod_version
            world.models[self._name] = self

        def name(self):
            return self._name

        def to_gazebo(self):
            return self

        def reset_joint_positions(self, pos, names):
            self.last_pos, self.last_names = np.array(pos, dtype=float), list(names)
            return True

        def reset_joint_velocities(self, vel, names):
            self.last_vel = np.array(vel, dtype=float)
            return True
    models_pkg = sys.modules['gym_os2r.models']
    models_pkg.monopod = mod('gym_os2r.models.monopod', Monopod=FakeMonopod,
                             get_model_file_from_name=lambda name: name)
    return FakeMonopod


class FakeWorld:
    def __init__(self):
        self.models = {}

    def model_names(self):
        return list(self.models)

    def to_gazebo(self):
        return self

    def remove_model(self, name):
        self.models.pop(name)
        return True

    def get_model(self, name):
        return self.models[name]


class FakeGazebo:
    def run(self, paused=False):
        return True


def main():
    gg._stub_modules()
    gg._load('gym_os2r.models.config', 'gym_os2r/models/config/__init__.py')
    gg._load('gym_os2r.rewards.rewards_utils', 'gym_os2r/rewards/rewards_utils.py')
    rw = gg._load('gym_os2r.rewards', 'gym_os2r/rewards/__init__.py')
    t_norm = gg._load('gym_os2r.tasks.monopod', 'gym_os2r/tasks/monopod.py')
    sys.modules['gym_os2r.tasks'].monopod = t_norm
    sys.modules['gym_os2r'].tasks = sys.modules['gym_os2r.tasks']
    gg._load('gym_os2r.utils.reset', 'gym_os2r/utils/reset.py')
    _extra_stubs()
    sys.modules['gym_os2r'].models = sys.modules['gym_os2r.models']
    for pkg in ('gym_os2r.randomizers',):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, *pkg.split('.'))]
        sys.modules[pkg] = m
    rnd = gg._load('gym_os2r.randomizers.monopod', 'gym_os2r/randomizers/monopod.py')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')       # `is not 'simple'` SyntaxWarning of the reference file
        nornd = gg._load('gym_os2r.randomizers.monopod_no_rand', 'gym_os2r/randomizers/monopod_no_rand.py')

    class Mixin(rnd.MonopodRandomizersMixin):
        # SDF regeneration (gym-ignition SDFRandomizer + lxml) is not on the reset-pose path
        def randomize_model_description(self, task, **kwargs):
            return None

        def randomize_ground_description(self, task, **kwargs):
            return None

    def make_task(mode, poses, reward='BalancingV1'):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            task = t_norm.MonopodTask(agent_rate=1000, task_mode=mode, reward_class=getattr(rw, reward),
                                      reset_positions=list(poses))
        aspace, ospace = task.create_spaces()
        task.action_space, task.observation_space = aspace, ospace
        task.world, task.model_name, task.model = FakeWorld(), None, None
        return task

    out = {}
    gz = FakeGazebo()
    # ---- MonopodEnvRandomizer path (randomizers/monopod.py:67-135); needs a yaw joint (:121)
    for mode in ('fixed_hip', 'free_hip'):
        for k, poses in enumerate([[p] for p in POSES] + [POSES]):
            task = make_task(mode, poses, 'BalancingV1')
            mix = Mixin()
            np.random.seed(1000 + 17 * k + (0 if mode == 'fixed_hip' else 500))
            rows, chosen = [], []
            for _ in range(N_SAMPLES):
                mix.randomize_task(task, gazebo=gz)
                m = task.world.get_model(task.model_name)
                assert m.last_names == task.joint_names and not m.last_vel.any()
                rows.append(m.last_pos)
                chosen.append(POSES.index(task.current_reset_orientation))
            key = f'rand/{mode}/' + (poses[0] if len(poses) == 1 else 'all')
            out[key] = np.array(rows, dtype=np.float32)
            if len(poses) > 1:
                out[key + '/chosen'] = np.array(chosen, dtype=np.int8)
        out[f'rand/{mode}/joint_names'] = np.array(task.joint_names)
    # ---- MonopodEnvNoRandomizer path (monopod_no_rand.py:26-98)
    for mode in ('fixed_hip', 'free_hip', 'fixed', 'fixed_hip_simple', 'simple'):
        for pose in POSES:
            task = make_task(mode, [pose], 'StraightV1' if mode == 'simple' else 'BalancingV1')
            obj = object.__new__(nornd.MonopodEnvNoRandomizer)
            np.random.seed(7)
            n = N_SAMPLES if mode == 'simple' else 3
            if mode == 'simple':
                task.observation_space.seed(11)
            rows = []
            for _ in range(n):
                obj.randomize_task(task, gazebo=gz)
                rows.append(task.world.get_model(task.model_name).last_pos)
            rows = np.array(rows)
            if mode != 'simple':
                assert (rows == rows[0]).all()        # the NoRandomizer pose is deterministic
                rows = rows[:1]
            out[f'norand/{mode}/{pose}'] = rows.astype(np.float32 if mode == 'simple' else np.float64)
            if mode == 'simple':
                break                                  # the pose name plays no role in `simple` mode (:84)
        out[f'norand/{mode}/joint_names'] = np.array(task.joint_names)
    path = os.path.join(ROOT, 'tests', 'golden', 'reset_poses.npz')
    np.savez_compressed(path, **out)
    print(f'wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB')
    for k in sorted(out):
        if out[k].dtype.kind == 'f':
            print(f'  {k:34s} {out[k].shape}  mean {np.round(out[k].mean(0), 4).tolist()}')


if __name__ == '__main__':
    main()
