#!/usr/bin/env python
"""Probe: the packed host step (numpy in / out, what `e2e` measures) with and without the two-launch split
(os2r_tuning.disable_host_split), fixed_hip 65 536 envs and free_hip 131 072 envs."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from gym_os2r_b200.runtimes.engine import Engine  # noqa: E402
from helpers import make_config  # noqa: E402


def run(mode, N, split):
    task, cm, cfg = make_config(mode, reward='BalancingV1' if mode == 'fixed_hip' else 'HoppingV1', randomize_params=True,
                                randomize_gravity=True, reset_randomized=True, auto_reset=True, max_episode_steps=100000,
                                pgs_tol=1e-6)
    eng = Engine(cm, cfg, N, seed=42, tuning=None if split else {'disable_host_split': 1})
    eng.reset()
    rng = np.random.RandomState(0)
    buf = eng.action_buffer
    acts = [rng.uniform(-1, 1, (N, 2)).astype(np.float32) for _ in range(8)]
    for i in range(600):
        buf[:] = acts[i % 8]
        eng.step_host_packed(buf)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(200):
            eng.step_host_packed(buf)
        best = min(best, (time.perf_counter() - t0) / 200)
    eng.close()
    return f'{mode} N={N} split={split}: {best * 1e6:.1f} us/step -> {N / best / 1e6:.1f} M env-steps/s end to end'


if __name__ == '__main__':
    for mode, N in (('free_hip', 131072), ('fixed_hip', 131072), ('fixed_hip', 262144)):
        for split in (False, True):
            print(run(mode, N, split), flush=True)
