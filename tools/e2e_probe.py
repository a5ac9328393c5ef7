#!/usr/bin/env python
"""Break down the host-buffer (e2e) step time."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_os2r_b200 import randomizers
from gym_os2r_b200.common import make_mp_envs
N = 65536
envs = make_mp_envs('Monopod-balance-v1', N, 42, randomizers.monopod.MonopodEnvRandomizer, task_mode='fixed_hip')
envs.reset()
eng = envs.runtime.engine
rng = np.random.RandomState(0)
acts = [rng.uniform(-1, 1, (N, 2)).astype(np.float32) for _ in range(8)]
for i in range(400): envs.step(acts[i % 8])   # past touchdown: some envs finish every step
def t(fn, n=200):
    for i in range(5): fn(i)
    t0 = time.perf_counter()
    for i in range(n): fn(i)
    return (time.perf_counter() - t0) / n * 1e6
print('VecEnv.step(numpy)            us:', t(lambda i: envs.step(acts[i % 8])))
print('engine.step_host_packed        us:', t(lambda i: eng.step_host_packed(acts[i % 8])))
import ctypes as C
_blk = torch.empty(4 * 1024 * 1024, dtype=torch.uint8).pin_memory().numpy(); _k = C.c_int32()
def one_call(i):
    a = acts[i % 8]
    eng.lib.os2r_step_host_packed(eng.handle, a.ctypes.data_as(C.c_void_p), _blk.ctypes.data_as(C.c_void_p), 1024, C.byref(_k))
print('C os2r_step_host_packed (one call) us:', t(one_call))
print('engine.step_host_packed again  us:', t(lambda i: eng.step_host_packed(acts[i % 8])))
if os.environ.get('E2E_SHORT'):
    sys.exit(0)
print('engine.step_host (obs,rew,done) us:', t(lambda i: eng.step_host(acts[i % 8])))
print('engine.step_host (+term,+info)  us:', t(lambda i: eng.step_host(acts[i % 8], True, True)))
a_dev = [torch.as_tensor(a, device='cuda') for a in acts]
def dev(i):
    eng.step(a_dev[i % 8]); torch.cuda.synchronize()
print('device step + sync            us:', t(dev))
pin_o = torch.empty((N, 8), dtype=torch.float32).pin_memory(); pin_a = torch.empty((N, 2), dtype=torch.float32).pin_memory()
def pinned(i):
    a_dev[0].copy_(pin_a, non_blocking=True); eng.step(a_dev[0]); pin_o.copy_(eng.obs, non_blocking=True); torch.cuda.synchronize()
print('pinned H2D + step + pinned D2H obs us:', t(pinned))
buf = np.empty((N, 8), np.float32)
print('np.empty((N,8)) + fill        us:', t(lambda i: np.empty((N, 8), np.float32).fill(0)))
print('memcpy 2MB np                 us:', t(lambda i: np.copyto(buf, pin_o.numpy())))
pin_acts = [torch.as_tensor(a).pin_memory().numpy() for a in acts]
print('engine.step_host_packed, pinned actions us:', t(lambda i: eng.step_host_packed(pin_acts[i % 8])))
blk = torch.empty(3 * 1024 * 1024, dtype=torch.uint8).pin_memory(); dblk = torch.empty(3 * 1024 * 1024, dtype=torch.uint8, device='cuda')
def d2h(i):
    blk.copy_(dblk, non_blocking=True); torch.cuda.synchronize()
print('pinned D2H 3 MB + sync        us:', t(d2h))
def h2d(i):
    a_dev[0].copy_(pin_a, non_blocking=True); torch.cuda.synchronize()
print('pinned H2D 512 KB + sync      us:', t(h2d))
print('memcpy 512KB np               us:', t(lambda i: np.copyto(pin_a.numpy(), acts[i % 8])))
