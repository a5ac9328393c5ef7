#!/usr/bin/env python
"""Probe: what one launch of the step kernel costs OUTSIDE the physics loop (lane sort, prologue loads, fp64 task
epilogue, stores) — the kernel timed with 0, 1, 2 and 10 physics iterations per env step. Appends to gpurun_out/kprobe.log."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kprobe import ROOT, timeit  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from gym_os2r_b200.runtimes.engine import Engine  # noqa: E402
from helpers import make_config  # noqa: E402


def run(substeps, mode='fixed_hip', N=65536, reward='BalancingV1', pre=300):
    task, cm, cfg = make_config(mode, reward=reward, randomize_params=True, randomize_gravity=True, reset_randomized=True,
                                auto_reset=True, max_episode_steps=100000, pgs_tol=1e-6, substeps=substeps)
    eng = Engine(cm, cfg, N, seed=42)
    eng.reset()
    g = torch.Generator(device='cuda')
    g.manual_seed(0)
    acts = [torch.rand((N, 2), device='cuda', generator=g) * 2 - 1 for _ in range(64)]
    for i in range(pre):
        eng.step(acts[i % 64])
    ms = min(timeit(lambda i: eng.step(acts[i % 64]), 200) for _ in range(3))
    eng.close()
    return f'{mode} {reward} N={N} substeps={substeps}: {ms * 1e3:.1f} us/step'


if __name__ == '__main__':
    out = open(os.path.join(ROOT, 'gpurun_out', 'kprobe.log'), 'a')
    for mode, reward in (('fixed_hip', 'BalancingV1'), ('fixed_hip', 'HoppingV1'), ('free_hip', 'HoppingV1')):
        for k in (0, 1, 2, 10):
            s = run(k, mode=mode, reward=reward, pre=300 if k else 20)
            print(s, flush=True)
            out.write(s + '\n')
