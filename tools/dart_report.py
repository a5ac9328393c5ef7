#!/usr/bin/env python
"""Tolerance report of this backend against a golden trajectory file recorded by tools/record_dart_golden.py
(BASELINE.json north_star: contact-free max |dq| <= 1e-4 rad and |dqd| <= 1e-3 rad/s over 1000 steps; with contact:
reward within 1e-3 relative while states agree, touchdown timing and episode return within a documented tolerance; done
flags bit-exact given matching states).

    python tools/dart_report.py tests/golden/dart_fixed_hip.npz [--json out.json]

Replays the file's action sequence through the fp32 engine from the same reset and prints: the error-vs-step curve
(max and median over envs at a handful of steps), the first step at which the contact-free bounds are exceeded, the
touchdown-step difference histogram (first step with a link in contact on either side), the return difference, and
reward / done mismatches over the steps where the states still agree."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def report(path, agree=(1e-4, 1e-3)):
    import torch
    from gym_os2r_b200.runtimes.engine import Engine
    from helpers import make_config
    g = np.load(path, allow_pickle=True)
    mode, reset = str(g['task_mode']), str(g['reset'])
    reward = {'Monopod-balance-v1': 'BalancingV1', 'Monopod-balance-v2': 'BalancingV2', 'Monopod-stand-v1': 'StandingV1',
              'Monopod-hop-v1': 'HoppingV1', 'Monopod-simple-v1': 'StraightV1'}.get(str(g['env']), 'BalancingV1')
    task, cm, cfg = make_config(mode, reward=reward, reset_positions=(reset,), pgs_tol=1e-6)
    acts, q_ref, qd_ref = g['actions'], g['q'], g['qd']                    # [T, E, 2], [T, E, nj], [T, E, nj]
    T, E = acts.shape[:2]
    n, nc = cm.n_dof, cm.struct.n_contacts
    order = [cm.dof_of(str(name)) for name in g['joint_names']]
    eng = Engine(cm, cfg, E, precision=32)
    eng.reset()
    dq, dv = np.zeros((T, E)), np.zeros((T, E))
    td_us, td_ref = np.full(E, -1), np.full(E, -1)
    ret_us, ret_ref = np.zeros(E), np.zeros(E)
    rew_bad = done_bad = 0
    has_contact = 'in_contact' in g.files and g['in_contact'].any()
    for t in range(T):
        obs, rew, done, _ = eng.step(torch.as_tensor(acts[t].astype(np.float32), device='cuda'))
        s = eng.get_state()
        dq[t] = np.abs(s[:, order] - q_ref[t]).max(1)
        dv[t] = np.abs(s[:, [n + d for d in order]] - qd_ref[t]).max(1)
        touching = (s[:, 3 * n:3 * n + 3 * nc:3] > 0).any(1)
        td_us = np.where((td_us < 0) & touching, t, td_us)
        if has_contact:
            td_ref = np.where((td_ref < 0) & g['in_contact'][t].astype(bool), t, td_ref)
        r_us = rew.cpu().numpy().astype(np.float64)
        ret_us += r_us; ret_ref += g['reward'][t]
        ok = (dq[t] < agree[0]) & (dv[t] < agree[1])
        rew_bad += int((np.abs(r_us - g['reward'][t])[ok] > 1e-3 * np.maximum(np.abs(g['reward'][t][ok]), 1e-6) + 1e-6).sum())
        done_bad += int((done.cpu().numpy().astype(bool)[ok] != g['done'][t].astype(bool)[ok]).sum())
    eng.close()
    over = np.nonzero((dq.max(1) > 1e-4) | (dv.max(1) > 1e-3))[0]
    marks = sorted(set(int(x) for x in np.unique(np.clip(np.round(np.geomspace(1, T, 12)).astype(int), 1, T)) - 1))
    both = (td_us >= 0) & (td_ref >= 0)
    hist = {}
    for d in np.abs(td_us - td_ref)[both]:
        hist[int(d)] = hist.get(int(d), 0) + 1
    return {'file': os.path.basename(path), 'source': str(g['source']) if 'source' in g.files else 'gym-ignition / DART',
            'task_mode': mode, 'reset': reset, 'steps': int(T), 'envs': int(E),
            'max_dq': float(dq.max()), 'max_dqd': float(dv.max()),
            'first_step_over_contact_free_bounds': int(over[0]) if len(over) else None,
            'first_touchdown_step_here': int(td_us[td_us >= 0].min()) if (td_us >= 0).any() else None,
            'curve': [{'step': m + 1, 'dq_max': float(dq[m].max()), 'dq_median': float(np.median(dq[m])),
                       'dqd_max': float(dv[m].max()), 'dqd_median': float(np.median(dv[m]))} for m in marks],
            'envs_landed_both': int(both.sum()), 'touchdown_diff_hist': hist,
            'return_here': float(ret_us.mean()), 'return_golden': float(ret_ref.mean()),
            'return_rel_diff': float(abs(ret_us.mean() - ret_ref.mean()) / max(1.0, abs(ret_ref.mean()))),
            'reward_mismatches': rew_bad, 'done_mismatches': done_bad}


def print_report(r):
    print(f"{r['file']} ({r['source']}): {r['task_mode']} / {r['reset']}, {r['envs']} envs x {r['steps']} steps")
    print(f"  max |dq| {r['max_dq']:.3e} rad, max |dqd| {r['max_dqd']:.3e} rad/s; first step over the contact-free bounds "
          f"(1e-4 rad / 1e-3 rad/s): {r['first_step_over_contact_free_bounds']}; first touchdown here at step {r['first_touchdown_step_here']}")
    for c in r['curve']:
        print(f"    step {c['step']:5d}: |dq| max {c['dq_max']:.2e} median {c['dq_median']:.2e}   |dqd| max {c['dqd_max']:.2e} median {c['dqd_median']:.2e}")
    print(f"  touchdown step difference histogram over {r['envs_landed_both']} envs that landed on both sides: {r['touchdown_diff_hist']}")
    print(f"  mean return {r['return_here']:.4f} here vs {r['return_golden']:.4f} golden (relative difference {r['return_rel_diff']:.2e}); "
          f"while states agree: {r['reward_mismatches']} reward mismatches (> 1e-3 relative), {r['done_mismatches']} done mismatches")


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('files', nargs='+')
    ap.add_argument('--json', default=None)
    a = ap.parse_args()
    out = [report(f) for f in a.files]
    for r in out:
        print_report(r)
    if a.json:
        with open(a.json, 'w') as f:
            json.dump(out, f, indent=1)
