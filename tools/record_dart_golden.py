#!/usr/bin/env python
"""Record DART golden trajectories from the REAL reference (gym-ignition + Ignition Gazebo + DART).

Cannot record the real thing in the build container or on the GPU box (gym-ignition is not installed there and there is
no network). Run it on a machine where `import gym_ignition, scenario, gym_os2r` (the reference) works:

    python tools/record_dart_golden.py --out tests/golden/dart_fixed_hip.npz --env Monopod-balance-v1 \
        --task-mode fixed_hip --steps 1000 --envs 8 --amplitude 0.1
    python tools/dart_report.py tests/golden/dart_fixed_hip.npz          # tolerance report against this backend

It replays the committed, seeded action sequence (sinusoidal, the same generator the tests use) through the reference
under MonopodEnvNoRandomizer, one env after the other through the plain gym API, and stores (q, qd, obs, reward, done,
in_contact) per step. `tests/` and tools/dart_report.py consume such a file when it is present and report "DART golden
absent - oracle-only parity" otherwise. Until a file recorded with this script is committed, every physics-parity
statement in this repo is against the fp64 oracle, NOT DART.

    python tools/record_dart_golden.py --selftest [--out gpurun_out/selftest_golden.npz]

runs the SAME recording loop against this repo's own runtime (the `gym_os2r` alias package resolves to gym_os2r_b200;
needs a GPU) and then the consumer (tools/dart_report.py): the file format, the recorder and the report are exercised
end to end, so that the day a machine with gym-ignition is available only the import changes.
"""
import argparse
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def actions(envs, steps, amplitude, seed=42):
    rng = np.random.RandomState(seed)
    phi = rng.uniform(0, 2 * np.pi, (envs, 2))
    f = np.array([1.0, 1.7])
    t = np.arange(steps)[:, None, None]
    return amplitude * np.sin(2 * np.pi * f * t / 1000.0 + phi[None])


def record(args, selftest):
    if selftest:
        sys.path.insert(0, ROOT)
    import gym_os2r
    from gym_os2r import randomizers
    from gym_os2r.common import make_env_from_id
    is_ours = hasattr(gym_os2r, '_impl') or 'b200' in (gym_os2r.__file__ or '')
    if selftest:
        assert is_ours, 'selftest records this repo\'s own runtime'
    else:
        import gym  # noqa: F401  (the real one)
        assert not is_ours, 'this must import the reference gym_os2r (run outside this repo, or use --selftest)'

    acts = actions(args.envs, args.steps, args.amplitude)
    rec = {k: [] for k in ('q', 'qd', 'obs', 'reward', 'done', 'in_contact')}
    joint_names = None
    for e in range(args.envs):
        make_env = functools.partial(make_env_from_id, env_id=args.env, task_mode=args.task_mode,
                                     reset_positions=[args.reset])
        env = randomizers.monopod_no_rand.MonopodEnvNoRandomizer(env=make_env)
        env.seed(42)
        env.reset()
        task = env.unwrapped.task
        joint_names = list(task.joint_names)
        q, qd, ob, rw, dn, ct = [], [], [], [], [], []
        for t in range(args.steps):
            o, r, d, _ = env.step(acts[t, e])
            q.append(task.model.joint_positions(joint_names))
            qd.append(task.model.joint_velocities(joint_names))
            ob.append(o); rw.append(r); dn.append(d)
            try:            # ScenarIO: Model.links_in_contact() -> names of the links touching anything
                ct.append(len(task.model.links_in_contact()) > 0)
            except Exception:
                ct.append(False)
        env.close()
        for k, v in zip(('q', 'qd', 'obs', 'reward', 'done', 'in_contact'), (q, qd, ob, rw, dn, ct)):
            rec[k].append(np.array(v))
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    np.savez_compressed(args.out, actions=acts, joint_names=np.array(joint_names), env=args.env,
                        task_mode=args.task_mode, reset=args.reset,
                        source='selftest: gym_os2r_b200 CUDA runtime' if selftest else 'gym-ignition / DART',
                        **{k: np.stack(v, 1) for k, v in rec.items()})
    print('wrote', args.out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--env', default='Monopod-balance-v1')
    ap.add_argument('--task-mode', default='fixed_hip')
    ap.add_argument('--reset', default='stand')
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--envs', type=int, default=8)
    ap.add_argument('--amplitude', type=float, default=0.1)
    ap.add_argument('--selftest', action='store_true')
    args = ap.parse_args()
    if args.selftest:
        args.out = args.out or os.path.join(ROOT, 'gpurun_out', 'selftest_golden.npz')
        args.steps, args.envs = min(args.steps, 300), min(args.envs, 4)
        record(args, selftest=True)
        sys.path.insert(0, os.path.join(ROOT, 'tools'))
        import dart_report
        rep = dart_report.report(args.out)
        dart_report.print_report(rep)
        # the recording came from this very runtime (single env through the gym API, batch replay through the engine):
        # identical arithmetic, so the report must show agreement to the last bit of the stored doubles
        assert rep['max_dq'] < 1e-9 and rep['max_dqd'] < 1e-6 and rep['touchdown_diff_hist'].get(0, 0) == rep['envs_landed_both'], rep
        assert rep['reward_mismatches'] == 0 and rep['done_mismatches'] == 0
        print('selftest ok: recorder -> file -> tolerance report run end to end')
        return
    if not args.out:
        ap.error('--out is required')
    record(args, selftest=False)


if __name__ == '__main__':
    main()
