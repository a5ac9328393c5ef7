import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_os2r_b200 import randomizers
from gym_os2r_b200.common import make_mp_envs
for env_id, mode in (('Monopod-balance-v1', 'fixed_hip'), ('Monopod-hop-v1', 'free_hip')):
    N = 65536
    envs = make_mp_envs(env_id, N, 7, randomizers.monopod.MonopodEnvRandomizer, task_mode=mode)
    envs.output = 'torch'
    envs.reset()
    eng = envs.runtime.engine
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    t0 = time.time()
    T = 30000
    for i in range(T):
        obs, rew, done, info = eng.step(torch.rand((N, 2), device='cuda', generator=g) * 2 - 1)
        if i % 5000 == 4999:
            torch.cuda.synchronize()
            st = eng.stats()
            print(mode, i + 1, 'steps', f'{time.time()-t0:.1f}s', {k: st[k] for k in ('episodes', 'done_task', 'done_timelimit', 'nonfinite_resets')},
                  'mean len', st['sum_length'] / max(st['episodes'], 1), 'obs finite', bool(torch.isfinite(obs).all()), 'max|obs|', float(obs.abs().max()), flush=True)
    s = eng.get_state()
    n = eng.model.n_dof
    print(mode, 'state finite', np.isfinite(s).all(), 'max|q|', np.abs(s[:, :n]).max(), 'max|qd|', np.abs(s[:, n:2*n]).max(), 'contact frac', (s[:, 3*n:3*n+3*eng.model.n_contacts:3] > 0).mean(0).round(3))
    envs.close()
