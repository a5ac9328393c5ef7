#!/usr/bin/env python
"""Pin the physics half of the oracle against ITSELF (regression fixture, not a parity claim: the reference holds no
trajectory). Writes tests/golden/oracle_physics_regression.json: per task mode one env dropped
from its reset pose under a fixed sinusoidal action, state snapshots at a few env steps, for the exact sweep mode
(pgs_tol = 0) and the production tolerance.

    python tools/gen_oracle_regression.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle
from helpers import make_config

CASES = [('simple', 'StraightV1', 'stand'), ('fixed', 'BalancingV1', 'stand'), ('fixed_hip', 'BalancingV1', 'stand'),
         ('fixed_hip', 'BalancingV1', 'lay'), ('free_hip', 'HoppingV1', 'ground')]
STEPS = (20, 60, 120, 200)


def trajectory(mode, reward, reset, tol):
    task, cm, cfg = make_config(mode, reward=reward, reset_positions=(reset,), pgs_tol=tol)
    orc = oracle.Oracle(cm.struct, cfg, 1, seed=5, nthreads=1)
    orc.reset()
    out = {}
    for t in range(1, max(STEPS) + 1):
        a = 0.6 * np.sin(2 * np.pi * np.array([3.0, 5.0]) * t / 1000.0 + np.array([0.3, 1.1]))
        obs, rew, done, _, _ = orc.step(a[None, :])
        if t in STEPS:
            out[str(t)] = {'state': orc.state[0].tolist(), 'obs': obs[0].tolist(), 'reward': float(rew[0])}
    return out


if __name__ == '__main__':
    data = {'steps': STEPS, 'cases': []}
    for mode, reward, reset in CASES:
        for tol in (0.0, 1e-6):
            data['cases'].append({'mode': mode, 'reward': reward, 'reset': reset, 'pgs_tol': tol,
                                  'snapshots': trajectory(mode, reward, reset, tol)})
    path = os.path.join(ROOT, 'tests', 'golden', 'oracle_physics_regression.json')
    with open(path, 'w') as f:
        json.dump(data, f)
    print('wrote', path, os.path.getsize(path), 'bytes')
