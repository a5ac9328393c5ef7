import sys, time, functools
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from gym_os2r import randomizers
from gym_os2r.common import make_env_from_id
env = randomizers.monopod.MonopodEnvRandomizer(env=functools.partial(make_env_from_id, env_id='Monopod-balance-v1', task_mode='fixed_hip'))
env.seed(1); env.reset()
rt = env.unwrapped; eng = rt.engine
a = np.array([[0.1,-0.2]], np.float32)
def t(fn, n=2000):
    for i in range(20): fn()
    t0=time.perf_counter()
    for i in range(n): fn()
    return (time.perf_counter()-t0)/n*1e6
print('env.step us', t(lambda: env.step([0.1,-0.2])))
print('engine.step_host us', t(lambda: eng.step_host(a, False, True)))
ad = torch.as_tensor(a, device='cuda')
def dev():
    eng.step(ad); torch.cuda.synchronize()
print('engine.step+sync us', t(dev))
print('_host_array us', t(lambda: eng._host_array('obs', (1,8), torch.float32)))
print('pool sizes', {k: len(v) for k,v in eng._host_pool.items()})
