#!/usr/bin/env python
"""Experiment: contact-free fp32 drift under the experimental precision flags (OS2R_FLAGS)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle
from gym_os2r_b200.runtimes.engine import Engine
from helpers import make_config

def run(mode, flags, seed, N=256, T=1000, A=0.1, reset='stand'):
    os.environ['OS2R_FLAGS'] = str(flags)
    task, cm, cfg = make_config(mode, reward='StraightV1' if mode == 'simple' else 'BalancingV1', reset_positions=(reset,))
    n = cm.n_dof
    eng = Engine(cm, cfg, N, seed=seed, precision=32)
    orc = oracle.Oracle(cm.struct, cfg, N, seed=seed, nthreads=16)
    eng.reset(); orc.reset(); orc.state[:] = eng.get_state()
    if mode != 'simple':
        st = eng.get_state(); st[:, cm.dof_of('planarizer_pitch_joint')] = 0.9; eng.set_state(st); orc.state[:] = eng.get_state()
    rng = np.random.RandomState(seed); phi = rng.uniform(0, 2*np.pi, (N, 2)); f = np.array([1.0, 1.7])
    out = []
    for t in range(T):
        a = (A*np.sin(2*np.pi*f*t/1000.0+phi)).astype(np.float32)
        eng.step(torch.as_tensor(a, device='cuda')); orc.step(a.astype(np.float64))
        if (t+1) % 250 == 0:
            sg = eng.get_state()
            dq = np.abs(sg[:, :n]-orc.state[:, :n]).max(1); dv = np.abs(sg[:, n:2*n]-orc.state[:, n:2*n]).max(1)
            out.append(f'{t+1}: dq max={dq.max():.1e} med={np.median(dq):.1e} dv max={dv.max():.1e}')
    eng.close()
    return ' | '.join(out)

if __name__ == '__main__':
    for mode in ('simple', 'fixed_hip'):
        for flags in (0, 1, 2, 3):
            for seed in (3, 4):
                print(mode, 'flags', flags, 'seed', seed, run(mode, flags, seed, T=1000 if mode=='simple' else 500), flush=True)
