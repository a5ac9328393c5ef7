"""Domain-randomising env wrapper — interface of gym_os2r/randomizers/monopod.py:27-385
(``MonopodRandomizersMixin`` / ``MonopodEnvRandomizer``).

In the reference this wrapper owns the reset: it removes and re-inserts the model with a freshly
sampled SDF (mass x U(0.8,1.2), joint friction U(0.01,0.05), damping x U(0.8,1.2), collision
mu = 0.33 x U(0.8,1.2); :182-215), draws gravity N(-9.8, 0.2) once (:56-61), and perturbs the reset
pose (:89-112). Here the same draws happen on the device, fused into the step kernel's auto-reset
(or the reset kernel); this wrapper only *configures* them on the wrapped ``CudaRuntime``.

Documented deviations: (a) draws come from one counter-based Philox stream per env keyed by
(seed, global env id, episode) — distribution-equal, not stream-equal, to numpy's global RNG;
(b) task modes without a yaw joint (``fixed``, ``simple``) work instead of raising ValueError
(:121); (c) the reference's ground-plane randomizer is built but never inserted (:338-347), so
only the link mu is randomised — reproduced.
"""
from typing import Callable

from .._gymshim import Wrapper


class MonopodRandomizersMixin:
    """Ranges of the SDF / physics randomisation; override ``randomization`` to change them."""

    randomization = dict(mass_lo=0.8, mass_hi=1.2, fric_lo=0.01, fric_hi=0.05, damp_lo=0.8, damp_hi=1.2,
                         mu_lo=0.8, mu_hi=1.2, mu_link=0.33, grav_mean=-9.8, grav_std=0.2)

    def __init__(self, randomize_physics_after_rollouts: int = 0):
        self.randomize_physics_after_rollouts = randomize_physics_after_rollouts

    def get_engine(self):
        return 'os2r-cuda'

    # The three hooks of the reference's randomizer ABCs. The work is done on the device; they
    # stay callable (no-ops returning what the reference returns) so subclasses that call super() work.
    def randomize_physics(self, task, **kwargs) -> None:
        return None

    def randomize_task(self, task, **kwargs) -> None:
        return None

    def randomize_model_description(self, task, **kwargs) -> str:
        return task.cfg.get_config(f'task_modes/{task.task_mode}/model')


class MonopodEnvRandomizer(Wrapper, MonopodRandomizersMixin):
    """``MonopodEnvRandomizer(env=make_env_callable, num_physics_rollouts=0)``."""

    def __init__(self, env: Callable, num_physics_rollouts: int = 0, **kwargs):
        MonopodRandomizersMixin.__init__(self, randomize_physics_after_rollouts=num_physics_rollouts)
        Wrapper.__init__(self, env() if callable(env) else env)
        if int(num_physics_rollouts) < 0:
            raise ValueError('num_physics_rollouts must be >= 0')
        # num_physics_rollouts = K > 0: gravity is drawn again at every K-th reset of an env (reference :36,56-61,371:
        # PhysicsRandomizer.physics_expired); 0 = drawn once at construction
        self.env.unwrapped.configure_randomization(reset_randomized=True, randomize_params=True,
                                                   randomize_gravity=True, randomization=dict(self.randomization),
                                                   gravity_redraw_resets=int(num_physics_rollouts))

    def get_state_info(self, state, actions):
        return self.env.unwrapped.task.get_state_info(state, actions)
