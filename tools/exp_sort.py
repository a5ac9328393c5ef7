#!/usr/bin/env python
"""Kernel A/B probe (kernel-only CUDA-event timings, hot L2, fresh random actions from a long pool).

    python tools/exp_sort.py <mode>        # results are appended to gpurun_out/exp_sort.log
      a  block width / sort margin / sweep tolerance sweep      d  sort-margin sweep at the production tolerance
      b  batch sizes and the other task modes                   e  steady state, from-reset average, sort off
      c  one library (OS2R_LIB=...) on the standard cases       g  strictly contact-free window (steps 5..60)
      h  BASELINE config 4 (free_hip) and the small models      l  small batches: OS2R_FORCE_LONE=0|1
Environment knobs read by libos2r.so at os2r_create: OS2R_FORCE_BLOCK=64|224, OS2R_SORT_MARGIN=<m> (-1 = sort off).
A variant library built with other -D flags is selected with OS2R_LIB=/path/to/lib.so (see DESIGN.md section 9 for
the experiments this was used for)."""
import os, sys
import torch
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


def run(mode='fixed_hip', N=65536, pre=1500, steps=200, iters=None, tol=None, env=None, scale=1.0, limit=100000):
    env = env or {}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        kw = dict(randomize_params=True, randomize_gravity=True, reset_randomized=True, auto_reset=True, max_episode_steps=limit)
        if iters is not None: kw['pgs_iters'] = iters
        if tol is not None: kw['pgs_tol'] = tol
        task, cm, cfg = make_config(mode, reward='BalancingV1' if mode != 'simple' else 'StraightV1', **kw)
        if os.environ.get('EXP_NC1'):      # I-cache footprint probe: keep only the hip proxy (lib built with -DOS2R_NC=1)
            cm.struct.n_contacts = 1
        eng = Engine(cm, cfg, N, seed=42)
        eng.reset()
        g = torch.Generator(device='cuda'); g.manual_seed(0)
        P = 256 if N <= 65536 else 64   # long pool: a short cycle gives every env a periodic torque with non-zero mean
        acts = [(torch.rand((N, 2), device='cuda', generator=g) * 2 - 1) * scale for _ in range(P)]
        for i in range(pre): eng.step(acts[i % P])
        ms = min(timeit(lambda i: eng.step(acts[i % P]), steps) for _ in range(3))
        st = eng.stats()
        lam = eng.get_state()[:, 3 * cm.n_dof:3 * cm.n_dof + 3 * cm.struct.n_contacts:3]
        info = eng.kernel_info()
        eng.close()
    finally:
        for k, v in old.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
    return (f'{mode} N={N} pre={pre} iters={iters} tol={tol} env={env} block={info.get("block_threads")}: {ms*1e3:.1f} us/step -> '
            f'{N/ms/1e3:.1f} M env-steps/s; contact frac={(lam>0).mean(0).round(3).tolist()} episodes={st["episodes"]}')


if __name__ == '__main__':
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    out = open(os.path.join(ROOT, 'gpurun_out', 'exp_sort.log'), 'a')
    def P(s):
        print(s, flush=True); out.write(s + '\n'); out.flush()
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if which in ('all', 'a'):
        P(run(env={'OS2R_FORCE_BLOCK': 64}))
        P(run(env={'OS2R_FORCE_BLOCK': 64, 'OS2R_SORT_MARGIN': -1.0}))
        P(run(env={'OS2R_FORCE_BLOCK': 224}))
        P(run(env={'OS2R_FORCE_BLOCK': 224, 'OS2R_SORT_MARGIN': -1.0}))     # every env classed "clear": sort = identity
        P(run(env={'OS2R_FORCE_BLOCK': 224, 'OS2R_SORT_MARGIN': 0.003}))
        P(run(env={'OS2R_FORCE_BLOCK': 224, 'OS2R_SORT_MARGIN': 0.03}))
        P(run(pre=0)); P(run(pre=0, env={'OS2R_FORCE_BLOCK': 64}))
        for it, tol in ((8, 1e-7), (16, 1e-7), (8, 1e-6), (16, 1e-8)):
            P(run(iters=it, tol=tol))
            P(run(iters=it, tol=tol, env={'OS2R_FORCE_BLOCK': 64}))
    if which == 'c':
        P('lib=' + os.environ.get('OS2R_LIB', 'default'))
        P(run()); P(run(env={'OS2R_FORCE_BLOCK': 64})); P(run(env={'OS2R_SORT_MARGIN': -1.0})); P(run(env={'OS2R_SORT_MARGIN': 0.003}))
        P(run(pre=0)); P(run(pre=0, env={'OS2R_FORCE_BLOCK': 64}))
        P(run(iters=8, tol=1e-7)); P(run(iters=8, tol=1e-7, env={'OS2R_SORT_MARGIN': -1.0})); P(run(iters=16, tol=1e-7))
    if which == 'd':
        for mg in (0.001, 0.002, 0.003, 0.005):
            P(run(iters=8, tol=1e-6, env={'OS2R_SORT_MARGIN': mg}))
        P(run(iters=8, tol=1e-6, env={'OS2R_SORT_MARGIN': -1.0}))
        P(run(iters=8, tol=1e-6, pre=0)); P(run(iters=8, tol=1e-6, pre=300)); P(run(iters=8, tol=1e-6, pre=5000))
        P(run(iters=6, tol=1e-6)); P(run(iters=8, tol=3e-6))
    if which == 'e':
        P('lib=' + os.environ.get('OS2R_LIB', 'default'))
        blk = os.environ.get('EXP_BLOCK')
        env = {'OS2R_FORCE_BLOCK': blk} if blk else {}
        P(run(iters=8, tol=1e-6, env=env)); P(run(iters=8, tol=1e-6, pre=0, env=env))
        P(run(iters=8, tol=1e-6, env=dict(env, OS2R_SORT_MARGIN=-1.0)))
    if which == 'g':     # strictly contact-free window: steps 5..65 after a reset, repeated
        P('lib=' + os.environ.get('OS2R_LIB', 'default'))
        kw = dict(randomize_params=True, randomize_gravity=True, reset_randomized=True, auto_reset=True, max_episode_steps=100000, pgs_iters=8, pgs_tol=1e-6)
        task, cm, cfg = make_config('fixed_hip', reward='BalancingV1', **kw)
        if os.environ.get('EXP_NC1'): cm.struct.n_contacts = 1
        N = 65536
        eng = Engine(cm, cfg, N, seed=42)
        g = torch.Generator(device='cuda'); g.manual_seed(0)
        acts = [(torch.rand((N, 2), device='cuda', generator=g) * 2 - 1) for _ in range(64)]
        best = 1e9
        for rep in range(6):
            eng.reset()
            for i in range(5): eng.step(acts[i])
            best = min(best, timeit(lambda i: eng.step(acts[5 + i]), 55))
        nc = cm.struct.n_contacts
        lam = eng.get_state()[:, 3 * cm.n_dof:3 * cm.n_dof + 3 * nc:3]
        P(f'contact-free window: {best*1e3:.1f} us/step; contact frac at the end {(lam>0).mean(0).round(4).tolist()}')
        eng.close()
    if which == 'h':     # BASELINE config 4: free_hip, production tolerance
        P(run(mode='free_hip', iters=8, tol=1e-6)); P(run(mode='free_hip', iters=8, tol=1e-6, N=131072, pre=800))
        P(run(mode='free_hip', iters=8, tol=1e-6, pre=0))
        P(run(mode='fixed', iters=8, tol=1e-6, pre=1000)); P(run(mode='simple', iters=8, tol=1e-6, pre=100))
    if which == 'l':     # small batches: build without an occupancy target (LONE) vs the 128-register build
        for N in (32, 1024, 4736, 9472):
            for lone in (0, 1):
                P(run(N=N, iters=8, tol=1e-6, pre=1500, steps=300, env={'OS2R_FORCE_LONE': lone}))
                P(run(N=N, iters=8, tol=1e-6, pre=0, steps=60, env={'OS2R_FORCE_LONE': lone}))
        for mode in ('free_hip', 'simple'):
            for lone in (0, 1):
                P(run(mode=mode, N=1024, iters=8, tol=1e-6, pre=300, steps=300, env={'OS2R_FORCE_LONE': lone}))
    if which == 'm':     # mid-size batches (2..4 narrow blocks per SM): does the many-register build still win?
        for N in (16384, 24576, 32768):
            for lone in (0, 1):
                P(run(N=N, iters=8, tol=1e-6, pre=1500, steps=300, env={'OS2R_FORCE_LONE': lone, 'OS2R_FORCE_BLOCK': 64}))
        P(run(N=32768, iters=8, tol=1e-6, pre=1500, steps=300, env={'OS2R_FORCE_BLOCK': 224}))
    if which in ('all', 'b'):
        for N in (9472, 16384, 33152, 131072):
            P(run(N=N)); P(run(N=N, env={'OS2R_FORCE_BLOCK': 64}))
        P(run(mode='free_hip')); P(run(mode='free_hip', env={'OS2R_FORCE_BLOCK': 64}))
        P(run(mode='simple', pre=100)); P(run(mode='fixed', pre=1000))
