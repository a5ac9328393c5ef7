#!/usr/bin/env python
"""Workload of tools/sanitize.sh (run against the CHECKED library, OS2R_LIB=.../libos2r_checked.so): every n_dof (task modes simple / fixed / fixed_hip / free_hip),
2 048 envs (+ a ragged 2 048 + 37 batch), 50 env steps with contacts, randomizers, TimeLimit auto-resets, the packed
host step, the wide lane-sorted blocks included (forced: the batch is far below their threshold)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from gym_os2r_b200.runtimes.engine import Engine  # noqa: E402
from helpers import make_config  # noqa: E402

import ctypes as C
import hashlib

from gym_os2r_b200 import _capi  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
digest = hashlib.sha256()
for mode in ('simple', 'fixed', 'fixed_hip', 'free_hip'):
    reward = 'StraightV1' if mode == 'simple' else 'BalancingV1'
    task, cm, cfg = make_config(mode, reward=reward, reset_positions=('ground', 'lay', 'stand'), auto_reset=True,
                                max_episode_steps=20, reset_randomized=mode in ('fixed_hip', 'free_hip'),
                                randomize_params=True, randomize_gravity=True, pgs_tol=1e-6)
    for N, tuning in ((2048, None), (2048 + 37, {'force_block': 224}), (333, {'force_block': 64})):
        eng = Engine(cm, cfg, N, seed=3, tuning=tuning)
        eng.reset()
        rng = np.random.RandomState(0)
        for t in range(steps):
            a = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
            if t % 2:
                eng.step_host_packed(a)
            else:
                eng.step(torch.as_tensor(a, device='cuda'))
        torch.cuda.synchronize()
        st = eng.get_state()
        assert np.isfinite(st).all()
        digest.update(st.tobytes())
        print(mode, N, tuning, eng.kernel_info(), 'episodes', eng.stats()['episodes'], flush=True)
        eng.close()
    if mode == 'fixed_hip':     # fp64 verification build once
        eng = Engine(cm, cfg, 512, seed=3, precision=64)
        eng.reset()
        for t in range(10):
            eng.step(torch.zeros((512, 2), device='cuda'))
        torch.cuda.synchronize()
        eng.close()
lib = _capi.load_library()
cnt = (C.c_uint64 * 8)()
_capi.check(lib.os2r_debug_counters(0, cnt, 0), lib)
names = ('source slot out of window', 'lane sort not a permutation', 'env index out of range', 'terminal records overflow',
         'cold-slot guard overwritten')
print('library', _capi.LIB_PATH, 'checked build' if cnt[7] else 'NOT a checked build')
for k, name in enumerate(names):
    print(f'  check[{k}] {name}: {cnt[k]}')
print('state digest', digest.hexdigest())
print('sanitize_run ok' if cnt[7] and not any(cnt[k] for k in range(5)) else 'sanitize_run FAILED')
