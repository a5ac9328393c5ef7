#!/usr/bin/env python
"""Kernel A/B probe: kernel-only CUDA-event timings of the step kernel (hot L2, fresh uniform actions from a long
pool) in the contact steady state and in the contact-free window after a reset.

    python tools/kprobe.py [case ...]      # results are appended to gpurun_out/kprobe.log
      std     fixed_hip 65 536 envs: steady state (pre-roll 1500) + contact-free window (steps 5..60 after a reset)
      cfg4    free_hip 65 536 / 131 072 envs (BASELINE config 4)
      small   N = 32 .. 32 768
      modes   fixed / simple
A variant library built with other -D flags is selected with OS2R_LIB=/path/to/lib.so; scheduling knobs travel in the
os2r_tuning struct (Engine(tuning=...)), never through the environment."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from gym_os2r_b200.runtimes.engine import Engine  # noqa: E402
from helpers import make_config  # noqa: E402


def timeit(fn, n):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def make(mode, N, tuning=None, iters=8, tol=1e-6, resets=('stand',)):
    kw = dict(randomize_params=True, randomize_gravity=True, reset_randomized=True, auto_reset=True,
              max_episode_steps=100000, pgs_iters=iters, pgs_tol=tol, reset_positions=resets)
    task, cm, cfg = make_config(mode, reward='BalancingV1' if mode != 'simple' else 'StraightV1', **kw)
    return cm, Engine(cm, cfg, N, seed=42, tuning=tuning)


def steady(mode='fixed_hip', N=65536, pre=1500, steps=200, tuning=None, **kw):
    cm, eng = make(mode, N, tuning, **kw)
    eng.reset()
    g = torch.Generator(device='cuda')
    g.manual_seed(0)
    P = 256 if N <= 65536 else 64   # long pool: a short cycle gives every env a periodic torque with non-zero mean
    acts = [torch.rand((N, 2), device='cuda', generator=g) * 2 - 1 for _ in range(P)]
    for i in range(pre):
        eng.step(acts[i % P])
    ms = min(timeit(lambda i: eng.step(acts[i % P]), steps) for _ in range(3))
    st = eng.stats()
    nc = cm.struct.n_contacts
    lam = eng.get_state()[:, 3 * cm.n_dof:3 * cm.n_dof + 3 * nc:3]
    info = eng.kernel_info()
    eng.close()
    return (f'{mode} N={N} pre={pre} tuning={tuning} {kw} block={info.get("block_threads")} regs={info.get("regs_per_thread")}: '
            f'{ms * 1e3:.1f} us/step -> {N / ms / 1e3:.1f} M env-steps/s; contact frac={(lam > 0).mean(0).round(3).tolist()} '
            f'episodes={st["episodes"]}')


def fresh(mode='fixed_hip', N=65536, tuning=None, **kw):
    """strictly contact-free window: steps 5..60 after a reset from `stand`, best of 6"""
    cm, eng = make(mode, N, tuning, **kw)
    g = torch.Generator(device='cuda')
    g.manual_seed(0)
    acts = [torch.rand((N, 2), device='cuda', generator=g) * 2 - 1 for _ in range(64)]
    best = 1e9
    for rep in range(6):
        eng.reset()
        for i in range(5):
            eng.step(acts[i])
        best = min(best, timeit(lambda i: eng.step(acts[5 + i]), 55))
    nc = cm.struct.n_contacts
    lam = eng.get_state()[:, 3 * cm.n_dof:3 * cm.n_dof + 3 * nc:3]
    eng.close()
    return f'{mode} N={N} tuning={tuning} contact-free window: {best * 1e3:.1f} us/step; contact frac at the end {(lam > 0).mean(0).round(4).tolist()}'


if __name__ == '__main__':
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    out = open(os.path.join(ROOT, 'gpurun_out', 'kprobe.log'), 'a')

    def P(s):
        print(s, flush=True)
        out.write(s + '\n')
        out.flush()
    cases = sys.argv[1:] or ['std']
    P('lib=' + os.environ.get('OS2R_LIB', 'default') + ' cases=' + ' '.join(cases))
    for which in cases:
        if which == 'std':
            P(steady()); P(fresh())
        if which == 'cfg4':
            P(steady(mode='free_hip')); P(steady(mode='free_hip', N=131072, pre=800)); P(fresh(mode='free_hip'))
        if which == 'small':
            for N in (32, 1024, 9472, 16384, 32768):
                P(steady(N=N, steps=300)); P(fresh(N=N))
        if which == 'modes':
            P(steady(mode='fixed', pre=1000)); P(fresh(mode='fixed')); P(steady(mode='simple', pre=100))
