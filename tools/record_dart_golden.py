#!/usr/bin/env python
"""Record DART golden trajectories from the REAL reference (gym-ignition + Ignition Gazebo + DART).

Cannot run in the build container or on the GPU box (gym-ignition is not installed there and there is no
network). Run it on a machine where `import gym_ignition, scenario, gym_os2r` (the reference) works:

    python tools/record_dart_golden.py --out tests/golden/dart_fixed_hip.npz --env Monopod-balance-v1 \
        --task-mode fixed_hip --steps 1000 --envs 8 --amplitude 0.1

It replays the committed, seeded action sequence (sinusoidal, the same generator tests use) through the reference
under MonopodEnvNoRandomizer and stores (q, qd, obs, reward, done) per step. `tests/` consume such a file when it is
present and report "DART golden absent - oracle-only parity" otherwise. Until a file recorded with this script is
committed, every physics-parity statement in this repo is against the fp64 oracle, NOT DART.
"""
import argparse
import functools

import numpy as np


def actions(envs, steps, amplitude, seed=42):
    rng = np.random.RandomState(seed)
    phi = rng.uniform(0, 2 * np.pi, (envs, 2))
    f = np.array([1.0, 1.7])
    t = np.arange(steps)[:, None, None]
    return amplitude * np.sin(2 * np.pi * f * t / 1000.0 + phi[None])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', required=True)
    ap.add_argument('--env', default='Monopod-balance-v1')
    ap.add_argument('--task-mode', default='fixed_hip')
    ap.add_argument('--reset', default='stand')
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--envs', type=int, default=8)
    ap.add_argument('--amplitude', type=float, default=0.1)
    args = ap.parse_args()

    import gym  # noqa: F401  (the real one)
    import gym_os2r  # noqa: F401  (the REFERENCE package, not this repo's alias)
    from gym_os2r import randomizers
    from gym_os2r.common import make_env_from_id
    assert 'b200' not in gym_os2r.__file__ and not hasattr(gym_os2r, '_impl'), 'this must import the reference gym_os2r'

    acts = actions(args.envs, args.steps, args.amplitude)
    rec = {k: [] for k in ('q', 'qd', 'obs', 'reward', 'done')}
    joint_names = None
    for e in range(args.envs):
        make_env = functools.partial(make_env_from_id, env_id=args.env, task_mode=args.task_mode,
                                     reset_positions=[args.reset])
        env = randomizers.monopod_no_rand.MonopodEnvNoRandomizer(env=make_env)
        env.seed(42)
        env.reset()
        task = env.unwrapped.task
        joint_names = list(task.joint_names)
        q, qd, ob, rw, dn = [], [], [], [], []
        for t in range(args.steps):
            o, r, d, _ = env.step(acts[t, e])
            q.append(task.model.joint_positions(joint_names))
            qd.append(task.model.joint_velocities(joint_names))
            ob.append(o); rw.append(r); dn.append(d)
        env.close()
        for k, v in zip(('q', 'qd', 'obs', 'reward', 'done'), (q, qd, ob, rw, dn)):
            rec[k].append(np.array(v))
    np.savez_compressed(args.out, actions=acts, joint_names=np.array(joint_names), env=args.env,
                        task_mode=args.task_mode, reset=args.reset, **{k: np.stack(v, 1) for k, v in rec.items()})
    print('wrote', args.out)


if __name__ == '__main__':
    main()
