#!/usr/bin/env python
"""Short run of the bench workload for ncu captures (same env construction as bench.py)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_os2r_b200 import randomizers
from gym_os2r_b200.common import make_mp_envs

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
mode = sys.argv[2] if len(sys.argv) > 2 else 'fixed_hip'
N = 65536
envs = make_mp_envs('Monopod-balance-v1' if mode != 'free_hip' else 'Monopod-hop-v1', N, 42,
                    randomizers.monopod.MonopodEnvRandomizer, task_mode=mode)
envs.output = 'torch'
envs.reset()
eng = envs.runtime.engine
g = torch.Generator(device='cuda'); g.manual_seed(0)
for i in range(steps):
    eng.step(torch.rand((N, 2), device='cuda', generator=g) * 2 - 1)   # fresh actions: a short cycle biases every env
torch.cuda.synchronize()
print('ok', eng.kernel_launches, eng.kernel_info())
