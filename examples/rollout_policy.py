"""BASELINE config 5: device-resident rollout — a torch MLP policy feeds the fused step kernel, obs / reward /
done never leave the GPU. One process per GPU under torchrun; episode statistics are all-reduced over NCCL."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import os
import time

import torch
import torch.distributed as dist

from gym_os2r import randomizers
from gym_os2r.common import make_mp_envs
from gym_os2r.common.distributed import reduce_stats

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
N = 65536
envs = make_mp_envs("Monopod-hop-v1", N, 42, randomizers.monopod.MonopodEnvRandomizer, start_idx=rank * N, device=local)
envs.output = 'torch'
obs = envs.reset()
policy = torch.nn.Sequential(torch.nn.Linear(obs.shape[1], 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                             torch.nn.Linear(64, 2), torch.nn.Tanh()).cuda()
steps, warmup = 500, 20
static_obs = obs.clone()
with torch.no_grad():
    for _ in range(warmup):                       # cuBLAS / allocator warm-up before capture
        static_obs.copy_(envs.step(policy(static_obs))[0])
    # one CUDA graph = policy forward + fused env step (+ obs copy-back): the launch-bound inner loop
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_obs.copy_(envs.step(policy(static_obs))[0])
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(steps):
        graph.replay()
    torch.cuda.synchronize(); dt = time.time() - t0
stats = reduce_stats(envs.runtime.engine.stats(), device=torch.device('cuda', local))
if rank == 0:
    print(f'{world} GPU(s): {world * N * steps / dt / 1e6:.1f} M env-steps/s incl. policy; episodes={stats["episodes"]:.0f} '
          f'mean return={stats["mean_return"]}')
envs.close()
if world > 1:
    dist.destroy_process_group()
