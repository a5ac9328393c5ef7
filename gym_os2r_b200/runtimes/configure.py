"""Turn (task class, task kwargs, runtime options) into the two C structs the step path needs."""
from typing import Tuple

from .. import _capi
from ..models import compiler
from ..tasks.monopod import build_task_cfg


def configure(task_cls, agent_rate: float = 1000, physics_rate: float = 10000, *,
              max_episode_steps: int = 0, auto_reset: bool = False, reset_randomized: bool = False,
              randomize_params: bool = False, randomize_gravity: bool = False, randomization: dict = None,
              gravity_redraw_resets: int = 0,
              pgs_iters: int = None, pgs_tol: float = None, pgs_joint_sweeps: int = None, substeps: int = None, **task_kwargs) -> Tuple[object, compiler.CompiledModel, _capi.TaskCfg]:
    """Create the task, its spaces, the compiled model tables and the device task configuration."""
    task = task_cls(agent_rate=agent_rate, **task_kwargs)
    task.create_spaces()
    physics = task.cfg.get_config('physics')
    physics['substeps'] = int(physics_rate / agent_rate) if substeps is None else int(substeps)  # gazebo_runtime.py:46-55
    physics['dt'] = 1.0 / physics_rate
    if pgs_iters is not None:
        physics['pgs_iters'] = int(pgs_iters)
    if pgs_tol is not None:
        physics['pgs_tol'] = float(pgs_tol)
    if pgs_joint_sweeps is not None:
        physics['pgs_joint_sweeps'] = int(pgs_joint_sweeps)
    model_name = task.cfg.get_config(f'task_modes/{task.task_mode}/model')
    compiled = compiler.compile_model(model_name, physics, max_torque=tuple(task.max_torques))
    cfg = build_task_cfg(task, compiled, max_episode_steps=max_episode_steps, auto_reset=auto_reset,
                         reset_randomized=reset_randomized, randomize_params=randomize_params,
                         randomize_gravity=randomize_gravity, randomization=randomization,
                         gravity_redraw_resets=gravity_redraw_resets)
    return task, compiled, cfg
