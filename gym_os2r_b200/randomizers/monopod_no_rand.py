"""Non-randomising env wrapper — interface of gym_os2r/randomizers/monopod_no_rand.py:14-101.

Every reset restores the nominal model and one of the task's ``reset_positions`` (chosen uniformly
when several are given, :60); leg angles come from the closed-form IK (utils/reset.py) or, in
``simple`` mode, from a sample of the observation space (:84). Fixes the reference's identity
comparison ``task.task_mode is not 'simple'`` (:69)."""
from typing import Callable

from .._gymshim import Wrapper


class MonopodEnvNoRandomizer(Wrapper):
    def __init__(self, env: Callable, **kwargs):
        Wrapper.__init__(self, env() if callable(env) else env)
        self.env.unwrapped.configure_randomization(reset_randomized=False, randomize_params=False,
                                                   randomize_gravity=False)

    def randomize_task(self, task, **kwargs) -> None:
        return None

    def get_state_info(self, state, actions):
        return self.env.unwrapped.task.get_state_info(state, actions)
