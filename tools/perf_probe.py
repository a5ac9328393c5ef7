#!/usr/bin/env python
"""Where does the step time go? (kernel-only timings with CUDA events; writes gpurun_out/perf.log)"""
import os, sys, warnings
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from gym_os2r_b200.runtimes.engine import Engine
from helpers import make_config

def timeit(fn, n):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def run(mode='fixed_hip', N=65536, randomize=True, limit=100000, pgs=None, action_scale=1.0, pre=0, steps=100, prec=32):
    kw = dict(randomize_params=randomize, randomize_gravity=randomize, reset_randomized=randomize, auto_reset=True, max_episode_steps=limit)
    if pgs is not None: kw['pgs_iters'] = pgs
    task, cm, cfg = make_config(mode, reward='BalancingV1' if mode != 'simple' else 'StraightV1', **kw)
    eng = Engine(cm, cfg, N, seed=42, precision=prec)
    eng.reset()
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    acts = [(torch.rand((N, 2), device='cuda', generator=g) * 2 - 1) * action_scale for _ in range(8)]
    for i in range(pre): eng.step(acts[i % 8])
    ms = timeit(lambda i: eng.step(acts[i % 8]), steps)
    st = eng.stats()
    lam = eng.get_state()[:, 3*cm.n_dof:3*cm.n_dof+9:3]
    t_reset = timeit(lambda i: eng.reset(), 5)
    eng.close()
    return f'{mode} N={N} prec={prec} pgs={pgs} scale={action_scale} pre={pre} limit={limit}: {ms*1e3:.1f} us/step -> {N/ms/1e3:.1f} M env-steps/s; episodes/step={st["episodes"]/(pre+steps):.1f}; envs with contact={(lam>0).any(1).mean():.2f}; full reset kernel={t_reset*1e3:.1f} us'

if __name__ == '__main__':
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    out = open(os.path.join(ROOT, 'gpurun_out', 'perf.log'), 'w')
    def P(s):
        print(s, flush=True); out.write(s + '\n'); out.flush()
    P(run(pre=0))                       # fresh: no resets yet
    P(run(pre=1500))                    # steady state with task resets
    P(run(pre=1500, limit=0))
    P(run(pre=0, action_scale=0.0))     # zero action: monopod falls and lies on the ground (contacts active)
    P(run(pre=300, action_scale=0.0))
    P(run(pre=300, action_scale=0.1))
    for pgs in (1, 4, 16):
        P(run(pre=300, action_scale=0.1, pgs=pgs))
    P(run(mode='free_hip', pre=0)); P(run(mode='free_hip', pre=1500))
    P(run(mode='simple', pre=0)); P(run(mode='fixed', pre=0))
    for N in (8192, 32768, 131072, 262144):
        P(run(N=N, pre=0))
    P(run(prec=64, pre=0))
