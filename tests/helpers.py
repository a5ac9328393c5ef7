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
