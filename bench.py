#!/usr/bin/env python
"""bench.py — monopod env-steps/s at 64K envs/GPU (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm (fp64 oracle on host cores)

A "step" is one env step (10 physics iterations of 1e-4 s + observation / reward / done /
auto-reset + randomiser draws on reset) of ALL envs of the rank: ONE fused kernel launch.
Workload (config.workload): BASELINE configs[2] in the shape of the reference's own CPU script
examples/multiprocessing_epochs.py — `Monopod-balance-v1` with task_mode='fixed_hip' (4 DoF, ground
contact, obs 8) under MonopodEnvRandomizer, uniform random actions, auto-reset — at 65 536 envs per GPU.
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

ENVS_PER_GPU = 65536
TASK_MODE = 'fixed_hip'
N_DOF, OBS_DIM = 4, 8


def algorithmic_flops(n, D, contact=True):
    """F(n) = 10*[(224n-259) + (205n-248) + 2n + 4n + C] + (12D + 60)."""
    C = 60 * n + 60 if contact else 0
    return 10 * ((224 * n - 259) + (205 * n - 248) + 2 * n + 4 * n + C) + (12 * D + 60)


def algorithmic_bytes(n, D):
    """fp32 state kept in registers across the 10 sub-steps."""
    read = 8 + 8 * n + 4 * (3 * n + 6) + 8
    write = 8 * n + 4 * D + 4 + 1 + 8
    return read + write


def make_cfg(randomize=True):
    from gym_os2r_b200 import rewards
    from gym_os2r_b200.runtimes.configure import configure
    from gym_os2r_b200.tasks.monopod import MonopodTask
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return configure(MonopodTask, task_mode=TASK_MODE, reward_class=rewards.BalancingV1,
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


def cpu_reference_run(steps, warmup, envs_per_core=128, cores=None):
    """Times the fp64 CPU oracle (oracle/os2r_oracle.c) on the host cores: the same workload on a
    bounded sample of `envs_per_core * cores` envs, one thread per core (the reference's shape is one
    Gazebo process per core, examples/multiprocessing_epochs.py:39)."""
    import numpy as np
    import oracle
    cores = cores or os.cpu_count() or 1
    task, cm, cfg = make_cfg()
    N = envs_per_core * cores
    orc = oracle.Oracle(cm.struct, cfg, N, seed=42, nthreads=cores)
    orc.reset()
    rng = np.random.RandomState(42)
    for i in range(warmup):
        orc.step(rng.uniform(-1, 1, (N, 2)))
    acts = [rng.uniform(-1, 1, (N, 2)) for _ in range(steps)]     # fresh actions every step, generated untimed
    t0 = time.perf_counter()
    for i in range(steps):
        orc.step(acts[i])
    dt = time.perf_counter() - t0
    return dict(value=N * steps / dt, ms_per_step=dt / steps * 1e3, n_envs=N, cores=cores,
                sample=f'{N} envs ({envs_per_core}/core) x {steps} env steps, fixed_hip + randomizers, fp64 C oracle, '
                       f'{cores} threads')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=50)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--envs-per-gpu', type=int, default=ENVS_PER_GPU)
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
    workload = (f'Monopod-balance-v1 task_mode={TASK_MODE} (4-DoF, ground contact, obs 8) + MonopodEnvRandomizer, '
                f'uniform random actions, auto-reset; {args.envs_per_gpu} envs/GPU; 1 step = 10 x 1e-4 s physics '
                f'+ obs/reward/done/reset')
    base = {'metric': 'monopod env-steps/sec at 64K envs/GPU', 'unit': 'env-steps/s', 'n_gpus': args.gpus,
            'steps': K, 'warmup': W, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'data': 'synthetic'}

    if args.impl == 'reference':
        # The reference's own path (Gazebo/DART) cannot run here; its CPU restatement is timed instead.
        if rank != 0:
            return
        r = cpu_reference_run(K, min(W, 5))
        line = dict(base, impl='reference', value=r['value'], ms_per_step=r['ms_per_step'], dtype='f64',
                    config={'workload': workload, 'sample_envs': r['n_envs'],
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
    from gym_os2r_b200.runtimes.engine import measure_fp32_peak

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    N = args.envs_per_gpu

    # public API: the call a user of the reference makes (examples/multiprocessing_epochs.py:46-48)
    envs = make_mp_envs('Monopod-balance-v1', N, 42, randomizers.monopod.MonopodEnvRandomizer,
                        start_idx=rank * N, task_mode=TASK_MODE, device=local_rank)
    envs.output = 'torch'
    rt = envs.runtime
    envs.reset()
    eng = rt.engine
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    # fresh uniform actions every step (as action_space.sample() in examples/multiprocessing_epochs.py:22-23);
    # a short cyclic pool would give every env a periodic torque with non-zero mean and spin the hip up.
    new_actions = lambda: torch.rand((N, 2), device=dev, generator=gen) * 2 - 1
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    fp32_peak, _ = measure_fp32_peak(local_rank)
    for i in range(max(W, 3)):
        eng.step(new_actions())
    torch.cuda.synchronize(dev)

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
    for i in range(K):
        flush.zero_()                      # L2 flush between timed iterations, outside the event pair
        a = new_actions()                  # action generation (torch) is outside the event pair too
        ev0[i].record()
        eng.step(a, want_terminal_obs=True, want_info=True)
        ev1[i].record()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - wall0
    sampler.stop_flag = True
    launches = eng.kernel_launches - launches0
    if world > 1:
        dist.barrier()
    per_step_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    total_ms = float(sum(per_step_ms))
    # hot-L2 variant: K back-to-back steps, no flush (state stays resident in the 126 MB L2)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pool = [new_actions() for _ in range(min(K, 512))]
    s0.record()
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

    # end-to-end: numpy actions in, numpy results out, through the VecEnv API -> os2r_step_host
    envs.output = 'numpy'
    K2 = max(3, min(K, 200))
    rng = np.random.RandomState(99 + rank)
    acts_h = [rng.uniform(-1, 1, (N, 2)).astype(np.float32) for _ in range(8)]
    for i in range(3):
        envs.step(acts_h[i % 8])
    e2e_s = 0.0
    if world > 1:
        dist.barrier()
    for i in range(K2):
        flush.zero_()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        obs_h, rew_h, done_h, _ = envs.step(acts_h[i % 8])
        e2e_s += time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    D = eng.obs_dim
    # actions in; out = ONE packed block: obs + reward + done + reset-id byte per env + the terminal-record prefix
    h2d, d2h = N * 2 * 4, int(next(iter(eng._packed_layouts.values())).total_bytes)

    # the path's only collective: reduce episode statistics over ranks (NCCL over NVLink)
    st = eng.stats()
    stats_vec = torch.tensor([st['episodes'], st['done_task'], st['done_timelimit'], st['nonfinite_resets'],
                              st['sum_return'], st['sum_length'], st['env_steps']], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats_vec, op=dist.ReduceOp.SUM)
    kinfo = eng.kernel_info()

    if rank == 0:
        value = world * N * K / (total_ms * 1e-3)
        F, B = algorithmic_flops(N_DOF, OBS_DIM), algorithmic_bytes(N_DOF, OBS_DIM)
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
        # (profiles/r1_step_kernel_steady.txt: dram__bytes_read.sum + dram__bytes_write.sum), else null.
        traffic = None
        try:
            tot = 0.0
            unit_scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            for ln in open(os.path.join(ROOT, 'profiles', 'r1_step_kernel_steady.txt')):
                if ln.startswith('dram__bytes_read.sum') or ln.startswith('dram__bytes_write.sum'):
                    _, val, unit = ln.split()
                    tot += float(val) * unit_scale[unit]
            traffic = tot or None
        except Exception:
            pass
        line = dict(base, value=value, ms_per_step=total_ms / K, dtype='f32',
                    config={'workload': workload, 'envs_per_gpu': N, 'n_dof': N_DOF, 'obs_dim': OBS_DIM,
                            'pgs_iters': int(rt._compiled.struct.pgs_iters), 'pgs_tol': float(rt._compiled.struct.pgs_tol),
                            'l2': 'flushed between timed steps (256 MiB memset outside each per-step CUDA-event pair); '
                                  'per-env state (26 MB) is otherwise L2-resident',
                            'hot_l2_value': world * N * K / (hot_ms * 1e-3), 'hot_l2_ms_per_step': hot_ms / K,
                            'wall_s_timed_region_incl_flush': wall, 'parallelism': f'env-sharded x{world}, no data-path collective'},
                    clocks=sampler.result(),
                    e2e={'value': world * N * K2 / e2e_s, 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d,
                         'd2h_bytes_per_step': d2h, 'steps': K2, 'api': 'make_mp_envs(...).step(numpy) -> os2r_step_host_packed (pageable numpy in, one pinned block out)'},
                    gpu_launches=int(launches),
                    roofline={'bound': 'fp32', 'achieved': achieved_tf, 'peak': fp32_peak, 'unit': 'TFLOP/s',
                              'frac': achieved_tf / fp32_peak if fp32_peak else None, 'traffic': traffic,
                              'traffic_note': 'bytes per launch, ncu capture in profiles/ (algorithmic: %d)' % (N * B),
                              'kernel': f'step_kernel<float,{N_DOF},3,{kinfo["block_threads"]}>', 'flop_per_env_step': F,
                              'peak_source': 'FFMA microbenchmark measured in this run (os2r_measure_fp32_peak); '
                                             'MEASURED_PEAKS.json has no fp32 entry',
                              'hbm': {'achieved': N * B / kern_s / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                                      'frac': N * B / kern_s / 1e9 / hbm_peak, 'bytes_per_env_step': B,
                                      'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback'},
                              **kinfo},
                    episode_stats={'episodes': stats_vec[0].item(), 'done_task': stats_vec[1].item(),
                                   'done_timelimit': stats_vec[2].item(), 'nonfinite_resets': stats_vec[3].item(),
                                   'mean_return': (stats_vec[4] / stats_vec[0]).item() if stats_vec[0] > 0 else None,
                                   'mean_length': (stats_vec[5] / stats_vec[0]).item() if stats_vec[0] > 0 else None})
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(steps=100, warmup=2, envs_per_core=128)
            line['cpu_baseline'] = {'value': r['value'], 'unit': 'env-steps/s', 'cores': r['cores'], 'kind': 'port',
                                    'sample': r['sample']}
        emit(line)
    envs.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
