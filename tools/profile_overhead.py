#!/usr/bin/env python
"""Short run of the step kernel with ZERO physics iterations (lane sort + prologue + task epilogue only) for an ncu
capture of the per-launch overhead: ncu ... -k regex:step_kernel --launch-skip 5 --launch-count 1 python tools/profile_overhead.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from gym_os2r_b200.runtimes.engine import Engine  # noqa: E402
from helpers import make_config  # noqa: E402

N = 65536
task, cm, cfg = make_config('fixed_hip', reward='BalancingV1', randomize_params=True, randomize_gravity=True,
                            reset_randomized=True, auto_reset=True, max_episode_steps=100000, pgs_tol=1e-6, substeps=0)
eng = Engine(cm, cfg, N, seed=42)
eng.reset()
g = torch.Generator(device='cuda')
g.manual_seed(0)
for i in range(10):
    eng.step(torch.rand((N, 2), device='cuda', generator=g) * 2 - 1)
torch.cuda.synchronize()
print('ok', eng.kernel_launches, eng.kernel_info())
