#!/usr/bin/env python
"""bench.py — monopod env-steps/s at 64K envs/GPU (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm (fp64 oracle on host cores)

A "step" is one env step (10 physics iterations of 1e-4 s + observation / reward / done /
auto-reset + randomiser draws on reset) of ALL envs of the rank: ONE fused kernel launch.
Workload (config.workload, default --config 3): BASELINE configs[2] in the shape of the reference's own CPU script
examples/multiprocessing_epochs.py — `Monopod-balance-v1` with task_mode='fixed_hip' (4 DoF, ground
contact, obs 8) under MonopodEnvRandomizer, uniform random actions, auto-reset — at 65 536 envs per GPU, timed in the
contact steady state the metric names ("with contact"): PREROLL_STEPS untimed steps from the reset come first, whatever
--warmup says. --config 4 / 5 select BASELINE's free-hopping 1 M-env and policy-rollout configurations.
Prints ONE JSON line (rank 0). See DESIGN.md section 6 for how every field is derived.
"""
import argparse
import json
import os
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json configs selectable with --config (3 = the configuration the metric is quoted on = default):
#   3  fixed-hip monopod with ground contact (examples/fixed_hip.py shape), 65 536 envs/GPU
#   4  free-hopping monopod with per-env randomizers, 131 072 envs/GPU (1 M envs over 8 GPUs)
#   5  full rollout loop: torch MLP policy -> fused step, CUDA-graph captured, 65 536 envs/GPU
CONFIGS = {
    3: dict(env_id='Monopod-balance-v1', task_mode='fixed_hip', reward='BalancingV1', n_dof=4, obs_dim=8, envs=65536,
            what='Monopod-balance-v1 task_mode=fixed_hip (4-DoF, ground contact, obs 8) + MonopodEnvRandomizer'),
    4: dict(env_id='Monopod-hop-v1', task_mode='free_hip', reward='HoppingV1', n_dof=5, obs_dim=10, envs=131072,
            what='Monopod-hop-v1 task_mode=free_hip (5-DoF free-hopping, ground contact, obs 10) + MonopodEnvRandomizer '
                 '(per-env mass / friction / damping / mu / gravity draws)'),
    5: dict(env_id='Monopod-hop-v1', task_mode='free_hip', reward='HoppingV1', n_dof=5, obs_dim=10, envs=65536,
            what='Monopod-hop-v1 (free_hip) + MonopodEnvRandomizer driven by a torch MLP policy 10-64-64-2 (tanh) on the '
                 'device, policy forward + fused env step captured in ONE CUDA graph'),
}
HEAD_START_FLUSHES = int(os.environ.get('OS2R_BENCH_HEAD', '32'))   # x 45 us of untimed GPU work in front of a timed window
PREROLL_STEPS = 1000     # untimed: from the `stand` reset the first touchdown happens around env step 90 and the
                         # collapse / bounce transient (more sweeps per iteration than later) lasts a few hundred more


def algorithmic_flops(n, D, contact=True):
    """F(n) = 10*[(224n-259) + (205n-248) + 2n + 4n + C] + (12D + 60)."""
    C = 60 * n + 60 if contact else 0
    return 10 * ((224 * n - 259) + (205 * n - 248) + 2 * n + 4 * n + C) + (12 * D + 60)


def algorithmic_bytes(n, D):
    """fp32 state kept in registers across the 10 sub-steps."""
    read = 8 + 8 * n + 4 * (3 * n + 6) + 8
    write = 8 * n + 4 * D + 4 + 1 + 8
    return read + write


def make_cfg(config=3, randomize=True):
    from gym_os2r_b200 import rewards
    from gym_os2r_b200.runtimes.configure import configure
    from gym_os2r_b200.tasks.monopod import MonopodTask
    c = CONFIGS[config]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return configure(MonopodTask, task_mode=c['task_mode'], reward_class=getattr(rewards, c['reward']),
                         reset_positions=['stand'], reset_randomized=randomize, randomize_params=randomize,
                         randomize_gravity=randomize, auto_reset=True, max_episode_steps=100_000)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.period, self.samples, self.reasons, self.stop_flag = period, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # NVML unavailable: report it rather than guess
            self.nv, self.h, self.err = None, None, str(e)

    def sample(self):
        if self.nv is None:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {'hw_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                     'hw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                     'sw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                     'sw_power_cap': getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4),
                     'hw_power_brake': getattr(nv, 'nvmlClocksThrottleReasonHwPowerBrakeSlowdown', 0x80)}
            for k, b in names.items():
                if bits & b:
                    self.reasons.add(k)
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(self.period)

    def result(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                    'note': getattr(self, 'err', 'no samples')}
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2], 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s)}


def cpu_reference_run(steps, warmup, config=3, envs_per_core=128, cores=None, preroll=PREROLL_STEPS):
    """Times the fp64 CPU oracle (oracle/os2r_oracle.c) on the host cores: the same workload on a bounded sample of
    `envs_per_core * cores` envs, one thread per core (the reference's shape is one Gazebo process per core,
    examples/multiprocessing_epochs.py:39), in the same regime as the GPU arm: `preroll` untimed steps from the reset
    bring the sample into the contact steady state before the clock starts."""
    import numpy as np
    import oracle
    cores = cores or os.cpu_count() or 1
    task, cm, cfg = make_cfg(config)
    N = envs_per_core * cores
    orc = oracle.Oracle(cm.struct, cfg, N, seed=42, nthreads=cores)
    orc.reset()
    rng = np.random.RandomState(42)
    for i in range(preroll + warmup):
        orc.step(rng.uniform(-1, 1, (N, 2)))
    acts = [rng.uniform(-1, 1, (N, 2)) for _ in range(steps)]     # fresh actions every step, generated untimed
    t0 = time.perf_counter()
    for i in range(steps):
        orc.step(acts[i])
    dt = time.perf_counter() - t0
    n, nc = cm.n_dof, cm.struct.n_contacts
    contact = (orc.state[:, 3 * n:3 * n + 3 * nc:3] > 0).mean(0).round(4).tolist()
    return dict(value=N * steps / dt, ms_per_step=dt / steps * 1e3, n_envs=N, cores=cores, contact_frac=contact,
                sample=f'{N} envs ({envs_per_core}/core) x {steps} env steps after {preroll} untimed pre-roll steps, '
                       f'{CONFIGS[config]["task_mode"]} + randomizers, fp64 C oracle, {cores} threads')


def contact_fractions(eng):
    """Share of envs whose contact proxy c carries a normal impulse right now (per proxy, and any)."""
    n, nc = eng.model.n_dof, eng.model.n_contacts
    lam = eng.get_state()[:, 3 * n:3 * n + 3 * nc:3] > 0
    return lam.mean(0).round(4).tolist(), float(lam.any(1).mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=50)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=3, choices=sorted(CONFIGS),
                    help='BASELINE.json config: 3 fixed_hip 64K envs/GPU (default, the metric), 4 free_hip 128K envs/GPU '
                         '+ randomizers, 5 policy rollout in a CUDA graph')
    ap.add_argument('--envs-per-gpu', type=int, default=0, help='override the config\'s env count per GPU')
    ap.add_argument('--preroll', type=int, default=PREROLL_STEPS,
                    help='untimed env steps from the reset before warm-up (contact steady state); independent of --warmup')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything a library prints there (e.g. the NCCL version banner) is sent
    # to stderr instead, and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + '\n').encode())

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    K, W = args.steps, max(args.warmup, 0)
    C = CONFIGS[args.config]
    N = args.envs_per_gpu or C['envs']
    N_DOF, OBS_DIM = C['n_dof'], C['obs_dim']
    workload = (f'BASELINE config {args.config}: {C["what"]}, uniform random actions, auto-reset; {N} envs/GPU; '
                f'1 step = 10 x 1e-4 s physics + obs/reward/done/reset; timed in the contact steady state '
                f'({args.preroll} untimed pre-roll steps from the reset)')
    base = {'metric': 'monopod env-steps/sec at 64K envs/GPU', 'unit': 'env-steps/s', 'n_gpus': args.gpus,
            'steps': K, 'warmup': W, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'data': 'synthetic'}

    if args.impl == 'reference':
        # The reference's own path (Gazebo/DART) cannot run here; its CPU restatement is timed instead.
        if rank != 0:
            return
        r = cpu_reference_run(K, min(W, 5), config=args.config if args.config != 5 else 4, preroll=args.preroll)
        line = dict(base, impl='reference', value=r['value'], ms_per_step=r['ms_per_step'], dtype='f64',
                    config={'workload': workload, 'sample_envs': r['n_envs'], 'contact_frac': r['contact_frac'],
                            'note': 'CPU restatement (fp64 C oracle), NOT Gazebo/DART: gym-ignition is not installable here'},
                    cpu_baseline={'value': r['value'], 'unit': 'env-steps/s', 'cores': r['cores'], 'kind': 'port',
                                  'sample': r['sample']},
                    e2e={'value': r['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                    gpu_launches=0)
        emit(line)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from gym_os2r_b200 import randomizers
    from gym_os2r_b200.common import make_mp_envs
    from gym_os2r_b200.common.distributed import bind_to_gpu_numa_node
    from gym_os2r_b200.runtimes.engine import measure_fp32_peak

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    # page-locked staging memory is first-touched by this process: sit on the GPU's NUMA node before allocating it
    numa = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    # public API: the call a user of the reference makes (examples/multiprocessing_epochs.py:46-48)
    kw = dict(task_mode=C['task_mode']) if args.config == 3 else {}
    envs = make_mp_envs(C['env_id'], N, 42, randomizers.monopod.MonopodEnvRandomizer, start_idx=rank * N,
                        device=local_rank, **kw)
    envs.output = 'torch'
    rt = envs.runtime
    obs0 = envs.reset()
    eng = rt.engine
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    # fresh uniform actions every step (as action_space.sample() in examples/multiprocessing_epochs.py:22-23);
    # a short cyclic pool would give every env a periodic torque with non-zero mean and spin the hip up.
    new_actions = lambda: torch.rand((N, 2), device=dev, generator=gen) * 2 - 1
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    fp32_peak, _ = measure_fp32_peak(local_rank)

    def head_start():
        # After a synchronize the launch queue is empty: the GPU reaches a step's first event record as soon as the
        # host issues it and then WAITS for the step kernel's launch to arrive, so the host's launch path (torch.rand,
        # ctypes, ~60 us) was billed to the first two event pairs of every timed window (108 and 92 us against 78:
        # OS2R_BENCH_TRACE=1). A few untimed flushes in front give the host its lead before the first pair instead
        # of after the second; the timed quantity stays the sum of the K per-step event pairs.
        for _ in range(HEAD_START_FLUSHES):
            flush.zero_()

    policy = graph = static_obs = None
    if args.config == 5:
        torch.manual_seed(7)
        policy = torch.nn.Sequential(torch.nn.Linear(OBS_DIM, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64),
                                     torch.nn.Tanh(), torch.nn.Linear(64, 2), torch.nn.Tanh()).to(dev)
        static_obs = obs0.clone()

        def one_step():
            with torch.no_grad():
                static_obs.copy_(eng.step(policy(static_obs))[0])
    else:
        def one_step():
            eng.step(new_actions(), want_terminal_obs=True, want_info=True)

    # untimed pre-roll into the contact steady state (random actions even for config 5: an untrained policy would
    # otherwise leave the monopods where the reset put them), then the W warm-up steps of the timed loop's own kind
    for i in range(args.preroll):
        eng.step(new_actions())
    if args.config == 5:
        static_obs.copy_(eng.obs)
        for i in range(3):
            one_step()                             # cuBLAS / allocator warm-up before capture
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            one_step()
        timed_step = graph.replay
    else:
        timed_step = one_step
    for i in range(max(W, 3)):
        flush.zero_()                          # warm-up steps run under the timed loop's own conditions (cold L2)
        timed_step()
    torch.cuda.synchronize(dev)
    contact0, any0 = contact_fractions(eng)
    eng.stats(clear=True)                          # episode statistics of the timed window only

    sampler = ClockSampler(local_rank)
    sampler.sample()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches0 = eng.kernel_launches
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    sampler.start()
    wall0 = time.perf_counter()
    head_start()
    if args.config == 5:
        for i in range(K):
            flush.zero_()
            ev0[i].record()
            timed_step()
            ev1[i].record()
    else:
        for i in range(K):
            flush.zero_()                      # L2 flush between timed iterations, outside the event pair
            a = new_actions()                  # action generation (torch) is outside the event pair too
            ev0[i].record()
            eng.step(a, want_terminal_obs=True, want_info=True)
            ev1[i].record()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - wall0
    sampler.stop_flag = True
    launches = (eng.kernel_launches - launches0) if args.config != 5 else K   # a graph replay launches the captured step kernel
    if world > 1:
        dist.barrier()
    per_step_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    total_ms = float(sum(per_step_ms))
    if os.environ.get('OS2R_BENCH_TRACE'):        # diagnosis only: the per-step device times of the timed window
        print('per_step_us', [round(x * 1e3, 1) for x in per_step_ms[:64]], file=sys.stderr)
    st = eng.stats()                               # timed window only
    contact1, any1 = contact_fractions(eng)
    # hot-L2 variant: K back-to-back steps, no flush (state stays resident in the 126 MB L2)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pool = [new_actions() for _ in range(min(K, 256))]
    for i in range(4):                             # untimed: the host's launch lead (see head_start) and a warm L2
        timed_step() if args.config == 5 else eng.step(pool[i % len(pool)])
    s0.record()
    if args.config == 5:
        for i in range(K):
            timed_step()
    else:
        for i in range(K):
            eng.step(pool[i % len(pool)])
    s1.record()
    torch.cuda.synchronize(dev)
    hot_ms = s0.elapsed_time(s1)
    t = torch.tensor([total_ms, hot_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)     # max over ranks, device-timed
    total_ms, hot_ms = float(t[0]), float(t[1])
    sampler.join(timeout=1.0)
    sampler.sample()

    # end-to-end through the VecEnv API with HOST buffers: actions wait in page-locked host memory (the env's own
    # `action_buffer`, where a host-side policy writes them), H2D + kernel + ONE D2H + sync inside the timed call,
    # numpy obs / reward / done / infos out
    envs.output = 'numpy'
    K2 = min(max(K, 50), 200)                  # enough steps for a stable mean even when the driver asks for K = 20
    rng = np.random.RandomState(99 + rank)
    acts_h = [rng.uniform(-1, 1, (N, 2)).astype(np.float32) for _ in range(8)]
    abuf = envs.action_buffer
    e2e_s = 0.0
    h2d = d2h = 0
    if args.config != 5:
        for i in range(12):                        # warm-up: the pool of page-locked result blocks grows to its steady size
            abuf[:] = acts_h[i % 8]                # (results are held exactly as in the timed loop)
            obs_h, rew_h, done_h, _ = envs.step(abuf)
        if world > 1:
            dist.barrier()
        for i in range(K2):
            flush.zero_()
            abuf[:] = acts_h[i % 8]               # the policy's output lands in the pinned buffer (untimed, as action
            torch.cuda.synchronize(dev)           # generation is in the device-timed loop)
            t0 = time.perf_counter()
            obs_h, rew_h, done_h, _ = envs.step(abuf)
            e2e_s += time.perf_counter() - t0
        h2d, d2h = N * 2 * 4, int(next(iter(eng._packed_layouts.values())).total_bytes)
        e2e_api = ('make_mp_envs(...).step(envs.action_buffer) -> os2r_step_host_packed (actions in the handle\'s page-locked '
                   'buffer, one pinned block out)')
    else:
        # device-resident rollout: nothing but the step's scalar result crosses PCIe (mean reward read back per step)
        for i in range(K2):
            flush.zero_()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            timed_step()
            r_mean = float(eng.reward.mean().item())
            e2e_s += time.perf_counter() - t0
        d2h = 4
        e2e_api = 'CUDA-graph replay (policy + step) + reward.mean().item() per step; observations never leave the device'
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    e2e_per_rank = [e2e_s / K2 * 1e3]
    if world > 1:
        tl = [torch.zeros_like(te) for _ in range(world)]
        dist.all_gather(tl, te)                       # per-rank figures travel with the line: ranks share the host's
        e2e_per_rank = [round(float(x[0]) / K2 * 1e3, 5) for x in tl]   # PCIe root ports and memory bandwidth
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])

    # secondary figure: the strictly contact-free window (steps 5..55 after a reset from `stand`, L2 flushed)
    free_ms = None
    if args.config != 5:
        envs.output = 'torch'
        envs.reset()
        for i in range(5):
            eng.step(new_actions())
        f0 = [torch.cuda.Event(enable_timing=True) for _ in range(50)]
        f1 = [torch.cuda.Event(enable_timing=True) for _ in range(50)]
        head_start()
        for i in range(50):
            flush.zero_()
            a = new_actions()
            f0[i].record()
            eng.step(a)
            f1[i].record()
        torch.cuda.synchronize(dev)
        free_ms = sum(a.elapsed_time(b) for a, b in zip(f0, f1)) / 50
        tf = torch.tensor([free_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        free_ms = float(tf[0])

    # the path's only collective: reduce episode statistics over ranks (NCCL over NVLink)
    stats_vec = torch.tensor([st['episodes'], st['done_task'], st['done_timelimit'], st['nonfinite_resets'],
                              st['sum_return'], st['sum_length'], st['env_steps']], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats_vec, op=dist.ReduceOp.SUM)
    kinfo = eng.kernel_info()

    if rank == 0:
        value = world * N * K / (total_ms * 1e-3)
        F = algorithmic_flops(N_DOF, OBS_DIM, contact=True)
        F_free = algorithmic_flops(N_DOF, OBS_DIM, contact=False)
        # SURVEY's C = 60n + 60 is the cost of ONE point contact: an env pays it per proxy that is pressing
        mean_contacts = float(sum(contact0) + sum(contact1)) / 2
        F_weighted = F_free + 10 * (60 * N_DOF + 60) * mean_contacts
        B = algorithmic_bytes(N_DOF, OBS_DIM)
        kern_s = total_ms * 1e-3 / K          # one launch per step: avg launch duration == ms_per_step
        achieved_tf = N * F / kern_s / 1e12
        peaks = {}
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        # DRAM traffic per launch of the step kernel from the committed `ncu --set full` capture of this workload
        # (dram__bytes_read.sum + dram__bytes_write.sum), else null.
        traffic, traffic_file = None, {3: 'r2b_step_kernel_steady.txt', 4: 'r2b_step_kernel_free_hip_steady.txt'}.get(args.config)
        try:
            tot = 0.0
            unit_scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            for ln in open(os.path.join(ROOT, 'profiles', traffic_file)):
                if ln.startswith('dram__bytes_read.sum') or ln.startswith('dram__bytes_write.sum'):
                    _, val, unit = ln.split()
                    tot += float(val) * unit_scale[unit]
            traffic = (tot * N / 65536) or None           # tools/profile_step.py captures 65 536 envs: scaled to this batch
        except Exception:
            pass
        nominal_peak = 148 * 128 * 2 * 1.965e9 / 1e12
        line = dict(base, value=value, ms_per_step=total_ms / K, dtype='f32',
                    config={'workload': workload, 'bench_config': args.config, 'envs_per_gpu': N, 'n_dof': N_DOF, 'obs_dim': OBS_DIM,
                            'pgs_iters': int(rt._compiled.struct.pgs_iters), 'pgs_tol': float(rt._compiled.struct.pgs_tol),
                            'preroll_steps': args.preroll,
                            'contact_frac': {'proxies': list(rt._compiled.contact_names),
                                             'at_start_of_timed_window': contact0, 'at_end_of_timed_window': contact1,
                                             'any_proxy': [round(any0, 4), round(any1, 4)]},
                            'l2': 'flushed between timed steps (256 MiB memset outside each per-step CUDA-event pair); '
                                  'per-env state is otherwise L2-resident',
                            'hot_l2_value': world * N * K / (hot_ms * 1e-3), 'hot_l2_ms_per_step': hot_ms / K,
                            'contact_free_value': (world * N / (free_ms * 1e-3)) if free_ms else None,
                            'contact_free_ms_per_step': free_ms,
                            'wall_s_timed_region_incl_flush': wall, 'numa_binding': numa,
                            'parallelism': f'env-sharded x{world}, no data-path collective'},
                    clocks=sampler.result(),
                    e2e={'value': world * N * K2 / e2e_s, 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d,
                         'd2h_bytes_per_step': d2h, 'steps': K2, 'ms_per_step': e2e_s / K2 * 1e3, 'ms_per_step_per_rank': e2e_per_rank,
                         'api': e2e_api},
                    gpu_launches=int(launches),
                    roofline={'bound': 'fp32', 'achieved': achieved_tf, 'peak': fp32_peak, 'unit': 'TFLOP/s',
                              'frac': achieved_tf / fp32_peak if fp32_peak else None, 'traffic': traffic,
                              'traffic_note': 'bytes per launch, ncu capture profiles/%s (algorithmic: %d)' % (traffic_file, N * B),
                              'kernel': f'step_kernel<float,{N_DOF},...,{kinfo["block_threads"]}>',
                              'flop_per_env_step': F,
                              'flop_note': 'SURVEY 8(d) F(n) with the contact term (the timed window is the contact steady '
                                           'state, see config.contact_frac); also given: the contact-free formula and the '
                                           'formula weighted by the measured number of pressing proxies per env',
                              'frac_contact_weighted': N * F_weighted / kern_s / 1e12 / fp32_peak if fp32_peak else None,
                              'flop_per_env_step_contact_weighted': F_weighted,
                              'contact_free': ({'flop_per_env_step': F_free, 'achieved': N * F_free / (free_ms * 1e-3) / 1e12,
                                                'frac': N * F_free / (free_ms * 1e-3) / 1e12 / fp32_peak} if free_ms and fp32_peak else None),
                              'peak_source': 'FFMA microbenchmark measured in this run (os2r_measure_fp32_peak); '
                                             'MEASURED_PEAKS.json has no fp32 entry',
                              'peak_nominal': nominal_peak, 'frac_of_nominal': achieved_tf / nominal_peak,
                              'hbm': {'achieved': N * B / kern_s / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                                      'frac': N * B / kern_s / 1e9 / hbm_peak, 'bytes_per_env_step': B,
                                      'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback'},
                              **kinfo},
                    episode_stats={'window': 'timed steps only', 'episodes': stats_vec[0].item(), 'done_task': stats_vec[1].item(),
                                   'done_timelimit': stats_vec[2].item(), 'nonfinite_resets': stats_vec[3].item(),
                                   'mean_return': (stats_vec[4] / stats_vec[0]).item() if stats_vec[0] > 0 else None,
                                   'mean_length': (stats_vec[5] / stats_vec[0]).item() if stats_vec[0] > 0 else None})
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(steps=60, warmup=2, config=args.config if args.config != 5 else 4, envs_per_core=128)
            line['cpu_baseline'] = {'value': r['value'], 'unit': 'env-steps/s', 'cores': r['cores'], 'kind': 'port',
                                    'sample': r['sample'], 'contact_frac': r['contact_frac']}
        emit(line)
    envs.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
