"""Shared helpers for the test-suite (configuration shortcuts)."""
import warnings

import numpy as np

from gym_os2r_b200 import rewards
from gym_os2r_b200.runtimes.configure import configure
from gym_os2r_b200.tasks import monopod, monopod_no_norm


def make_config(task_mode='fixed_hip', variant='norm', reward='BalancingV1', reset_positions=('stand',), **opts):
    """Task + compiled model + device task config. Unless a test asks otherwise the sweeps run in the exact mode
    (pgs_tol = 0: a fixed sweep count, early exit only at a bit-exact fixed point) so that the fp64 kernel and the
    oracle can be compared to 1e-11; the production default (settings.yaml physics/pgs_tol) is covered separately."""
    opts.setdefault('pgs_tol', 0.0)
    cls = monopod.MonopodTask if variant == 'norm' else monopod_no_norm.MonopodTask
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return configure(cls, task_mode=task_mode, reward_class=getattr(rewards, reward),
                         reset_positions=list(reset_positions), **opts)


def chain_state(task, compiled, q_joint_order, v_joint_order):
    """Re-order joint_names-ordered vectors (YAML order) into chain order."""
    n = compiled.n_dof
    q, v = np.zeros(n), np.zeros(n)
    for j, name in enumerate(task.joint_names):
        q[compiled.dof_of(name)] = q_joint_order[j]
        v[compiled.dof_of(name)] = v_joint_order[j]
    return q, v


def steady_state_parity(N=65536, sample=2048, preroll=400, horizons=(1, 20), mode='fixed_hip', seed=3, nthreads=16,
                        tight=(1e-5, 1e-2)):
    """Shipped configuration (fp32 kernel, production sweep tolerance, full batch) against the fp64 oracle in the contact
    steady state. Returns a report dict: per horizon (env steps after the hand-over) median / 95 % / max of the per-env
    max-abs joint position / velocity error, the fraction of envs over the `tight` bound, done-flag mismatches and the
    worst relative reward error among the envs whose states still agree (north_star contact criteria)."""
    import torch
    import oracle
    from gym_os2r_b200.runtimes.engine import Engine
    reward = 'HoppingV1' if mode == 'free_hip' else 'BalancingV1'
    task, cm, cfg = make_config(mode, reward=reward, auto_reset=True, max_episode_steps=100000,
                                reset_randomized=True, randomize_params=True, randomize_gravity=True, pgs_tol=1e-6)
    n, nc = cm.n_dof, cm.struct.n_contacts
    g = torch.Generator(device='cuda')
    g.manual_seed(5)
    eng = Engine(cm, cfg, N, seed=seed)
    info = eng.kernel_info()
    eng.reset()
    for _ in range(preroll):
        eng.step(torch.rand((N, 2), device='cuda', generator=g) * 2 - 1)
    S0, P0 = eng.get_state(), eng.get_params()
    steps0, ret0 = eng.get_episode()
    pick = np.sort(np.random.RandomState(0).choice(N, sample, replace=False))
    orc = oracle.Oracle(cm.struct, cfg, sample, seed=seed, nthreads=nthreads)
    orc.reset()
    orc.state[:] = S0[pick]
    orc.params[:] = P0[pick]
    orc.steps[:] = steps0[pick]
    orc.ret[:] = ret0[pick]
    rep = {'mode': mode, 'n_envs': N, 'sample': sample, 'preroll': preroll, 'block_threads': info['block_threads'],
           'regs_per_thread': info['regs_per_thread'], 'tight': list(tight),
           'contact_frac': (S0[:, 3 * n:3 * n + 3 * nc:3] > 0).mean(0).round(4).tolist(), 'horizons': {}}
    reset_seen = np.zeros(sample, dtype=bool)       # an env that reset on either side draws from different episodes
    t = 0
    for h in sorted(horizons):
        while t < h:
            a = torch.rand((N, 2), device='cuda', generator=g) * 2 - 1
            obs, rew, done, _ = eng.step(a)
            o_o, r_o, d_o, _, _ = orc.step(a.cpu().numpy()[pick].astype(np.float64))
            d_g = done.cpu().numpy().astype(bool)[pick]
            t += 1
            last = (obs.cpu().numpy()[pick], rew.cpu().numpy()[pick], d_g, o_o, r_o, d_o)
            reset_seen |= d_g | d_o
        sg = eng.get_state()[pick]
        keep = ~reset_seen
        eq = np.abs(sg[:, :n] - orc.state[:, :n]).max(1)[keep]
        ev = np.abs(sg[:, n:2 * n] - orc.state[:, n:2 * n]).max(1)[keep]
        agree = (eq < tight[0]) & (ev < tight[1])
        obs_g, rew_g, d_g, o_o, r_o, d_o = (x[keep] for x in last)
        rel = np.abs(rew_g - r_o) / np.maximum(np.abs(r_o), 1e-6)
        rep['horizons'][h] = {
            'envs_compared': int(keep.sum()),
            'dq_median': float(np.median(eq)), 'dq_p95': float(np.quantile(eq, 0.95)), 'dq_max': float(eq.max()),
            'dv_median': float(np.median(ev)), 'dv_p95': float(np.quantile(ev, 0.95)), 'dv_max': float(ev.max()),
            'envs_over_tight': int((~agree).sum()), 'frac_over_tight': float((~agree).mean()),
            'done_mismatch_where_states_agree': int((d_g[agree] != d_o[agree]).sum()),
            'reward_rel_err_where_states_agree': float(rel[agree].max()) if agree.any() else 0.0,
            'obs_err_where_states_agree': float(np.abs(obs_g[agree] - o_o[agree]).max()) if agree.any() else 0.0}
    eng.close()
    return rep
