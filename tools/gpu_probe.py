#!/usr/bin/env python
"""Exploratory GPU run: kernel vs oracle errors and a first throughput number (writes gpurun_out/probe.log)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import oracle  # noqa: E402
from gym_os2r_b200.runtimes.engine import Engine, measure_fp32_peak  # noqa: E402
from helpers import make_config  # noqa: E402


def random_state(task, cm, cfg, N, rng, contact=False):
    m = cm.struct
    n = m.n_dof
    W = 2 * n + (n + 3 * m.n_contacts) + 2
    st = np.zeros((N, W))
    st[:, :n] = rng.uniform(-0.6, 0.6, (N, n))
    if 'planarizer_pitch_joint' in cm.joint_names:
        st[:, cm.dof_of('planarizer_pitch_joint')] = rng.uniform(-0.04, 0.12, N) if contact else rng.uniform(0.25, 0.5, N)
    if contact:
        st[:, cm.dof_of('hip_joint')] = rng.uniform(0.2, 1.2, N)
        st[:, cm.dof_of('knee_joint')] = rng.uniform(-2.4, -0.4, N)
    st[:, n:2 * n] = rng.normal(0, 1.0, (N, n))
    return st


def compare(mode, N=256, steps=(1, 20), contact=False, randomize=False, seed=1):
    task, cm, cfg = make_config(mode, reward='StraightV1' if mode == 'simple' else 'BalancingV1',
                                randomize_params=randomize, randomize_gravity=randomize)
    m = cm.struct
    n = m.n_dof
    rng = np.random.RandomState(seed)
    st0 = random_state(task, cm, cfg, N, rng, contact)
    out = {}
    for prec in (64, 32):
        eng = Engine(cm, cfg, N, seed=seed, precision=prec)
        orc = oracle.Oracle(m, cfg, N, seed=seed, nthreads=8)
        if randomize:
            eng.reset(); orc.reset()
            par = eng.get_params()
            out[f'params_err_{prec}'] = np.abs(par - orc.params).max()
            orc.params[:] = par
        eng.set_state(st0)
        orc.state[:] = eng.get_state()
        out[f'setstate_roundtrip_{prec}'] = np.abs(orc.state - st0).max()
        r2 = np.random.RandomState(seed + 7)
        k = 0
        for target in steps:
            while k < target:
                a = r2.uniform(-1, 1, (N, 2)).astype(np.float32)
                o_g, r_g, d_g, _ = eng.step(torch.as_tensor(a, device='cuda'))
                o_o, r_o, d_o, _, _ = orc.step(a.astype(np.float64))
                k += 1
            sg = eng.get_state()
            dq = np.abs(sg[:, :n] - orc.state[:, :n]).max()
            dv = np.abs(sg[:, n:2 * n] - orc.state[:, n:2 * n]).max()
            dl = np.abs(sg[:, 2 * n:-2] - orc.state[:, 2 * n:-2]).max()
            dobs = np.abs(o_g.cpu().numpy() - o_o).max()
            ncontact = (orc.state[:, 3 * n:3 * n + 9:3] > 0).sum()
            out[f'p{prec}_k{target}'] = f'dq={dq:.2e} dv={dv:.2e} dlam={dl:.2e} dobs={dobs:.2e} done_eq={np.array_equal(d_g.cpu().numpy().astype(bool), d_o)} contacts={ncontact}'
        eng.close()
    return out


def trajectory(mode, N=128, T=1000, A=0.1, reset='stand', prec=32):
    task, cm, cfg = make_config(mode, reward='StraightV1' if mode == 'simple' else 'BalancingV1', reset_positions=(reset,))
    m = cm.struct
    n = m.n_dof
    eng = Engine(cm, cfg, N, seed=3, precision=prec)
    orc = oracle.Oracle(m, cfg, N, seed=3, nthreads=8)
    eng.reset(); orc.reset()
    orc.state[:] = eng.get_state()
    rng = np.random.RandomState(42)
    phi = rng.uniform(0, 2 * np.pi, (N, 2))
    f = np.array([1.0, 1.7])
    worst = (0, 0)
    curve = []
    for t in range(T):
        a = (A * np.sin(2 * np.pi * f * t / 1000.0 + phi)).astype(np.float32)
        eng.step(torch.as_tensor(a, device='cuda'))
        orc.step(a.astype(np.float64))
        if (t + 1) % 100 == 0:
            sg = eng.get_state()
            dq = np.abs(sg[:, :n] - orc.state[:, :n]).max()
            dv = np.abs(sg[:, n:2 * n] - orc.state[:, n:2 * n]).max()
            curve.append((t + 1, dq, dv))
    eng.close()
    return curve


def throughput(mode='fixed_hip', N=65536, steps=50, warmup=5, prec=32, randomize=True, pgs=None):
    kw = dict(randomize_params=randomize, randomize_gravity=randomize, reset_randomized=randomize, auto_reset=True,
              max_episode_steps=100000)
    if pgs is not None:
        kw['pgs_iters'] = pgs
    task, cm, cfg = make_config(mode, reward='BalancingV1', **kw)
    eng = Engine(cm, cfg, N, seed=42, precision=prec)
    eng.reset()
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    acts = [torch.rand((N, 2), device='cuda', generator=g) * 2 - 1 for _ in range(8)]
    for i in range(warmup):
        eng.step(acts[i % 8])
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps):
        eng.step(acts[i % 8])
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    info = eng.kernel_info()
    st = eng.stats()
    eng.close()
    return dict(mode=mode, N=N, prec=prec, ms_per_step=ms, env_steps_per_s=N / ms * 1e3, **info, episodes=st['episodes'])


if __name__ == '__main__':
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    log = open(os.path.join(ROOT, 'gpurun_out', 'probe.log'), 'w')

    def P(*a):
        s = ' '.join(str(x) for x in a)
        print(s); log.write(s + '\n'); log.flush()

    P(torch.cuda.get_device_name(0))
    P('fp32 peak TFLOP/s, MHz:', measure_fp32_peak(0))
    for mode in ('simple', 'fixed', 'fixed_hip', 'free_hip'):
        for contact in (False, True):
            if mode == 'simple' and contact:
                continue
            t = time.time()
            res = compare(mode, contact=contact, randomize=(mode == 'fixed_hip'))
            P(f'== compare {mode} contact={contact} ({time.time() - t:.1f}s)')
            for k, v in res.items():
                P('   ', k, v)
    for mode, reset in (('simple', 'stand'), ('fixed', 'float')):
        for prec in (64, 32):
            c = trajectory(mode, reset=reset, prec=prec, T=1000 if mode == 'simple' else 300)
            P(f'== trajectory {mode} prec={prec}:', ' '.join(f'[{t}: dq={dq:.1e} dv={dv:.1e}]' for t, dq, dv in c))
    for mode in ('fixed_hip', 'free_hip', 'simple'):
        for prec in (32, 64):
            P('== throughput', throughput(mode, prec=prec))
    for pgs in (2, 4, 16):
        P('== throughput pgs', pgs, throughput('fixed_hip', pgs=pgs))
    for N in (4096, 16384, 131072, 262144):
        P('== throughput N', throughput('fixed_hip', N=N))
