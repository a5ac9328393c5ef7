"""VecEnv facade over one batched ``CudaRuntime`` — API of gym_os2r/common/vec_env/vec_env.py:32-245
and subproc_vec_env.py:52-229 (``reset`` / ``step_async`` / ``step_wait`` / ``step`` /
``get_state_info`` / ``seed`` / ``close`` / ``get_attr`` / ``set_attr`` / ``env_method``).

``step_async`` enqueues the fused kernel on the current CUDA stream, ``step_wait`` synchronises and
hands back the results. ``output='numpy'`` (default, what reference scripts expect) returns host
arrays ``(obs[N,D] float32, rew[N], done[N] bool, infos)``; ``output='torch'`` keeps everything on the
device for a GPU-resident training loop. Auto-reset semantics follow the reference worker
(subproc_vec_env.py:14-21): on done the returned observation is the reset observation and the
terminal one is in ``infos``; reward/done are those of the terminal step."""
from collections.abc import Sequence

import numpy as np
import torch


class LazyInfos(Sequence):
    """The per-env ``info`` dicts of a step, built on access: materialising 65 536 dicts per step
    would cost more than the physics. ``infos[i]`` / iteration / ``len`` behave like the reference's tuple.

    The terminal observations are held sparsely — one row per env that finished an episode in this step
    (``terminal_indices`` / ``terminal_observations`` give them in bulk; ``infos[i]['terminal_observation']`` looks
    the row up) — because a dense [N, D] copy of them would double the device-to-host traffic of a step."""

    def __init__(self, names, reset_ids, done, term_idx, term_cause, term_obs):
        self._names, self._rid, self._done = names, reset_ids, done
        self.terminal_indices, self.terminal_causes, self.terminal_observations = term_idx, term_cause, term_obs
        self._row_of = None

    @classmethod
    def from_dense(cls, names, reset_ids, done, terminal_obs, cause):
        idx = np.flatnonzero(done)
        rows = terminal_obs[idx] if terminal_obs is not None else None
        return cls(names, reset_ids, done, idx, np.asarray(cause)[idx], rows)

    def __len__(self):
        return len(self._rid)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        d = {'reset_orientation': self._names[int(self._rid[i])]}
        if self._done[i]:
            if self._row_of is None:
                self._row_of = {int(e): k for k, e in enumerate(self.terminal_indices)}
            k = self._row_of[int(i)]
            if self.terminal_observations is not None:
                d['terminal_observation'] = self.terminal_observations[k]
            if (int(self.terminal_causes[k]) & 3) == 2:   # ended by the TimeLimit only (gym: truncated = not done-by-task)
                d['TimeLimit.truncated'] = True
        return d


class CudaVecEnv:
    def __init__(self, env, output: str = 'numpy'):
        assert output in ('numpy', 'torch')
        self.env = env
        self.runtime = env.unwrapped
        self.num_envs = self.runtime.num_envs
        self.observation_space = env.observation_space
        self.action_space = env.action_space
        self.output = output
        self.waiting = False
        self.closed = False
        self._pending = None

    @property
    def action_buffer(self) -> np.ndarray:
        """Page-locked ``[num_envs, 2]`` float32 array owned by the engine: write the actions here and pass THIS array to
        ``step`` / ``step_async`` and the host-to-device copy reads it directly (no staging memcpy). Do not write to it
        between ``step_async`` and ``step_wait``."""
        return self.runtime.engine.action_buffer

    def _out(self, t):
        return t.detach().cpu().numpy() if self.output == 'numpy' else t

    def reset(self):
        obs = self.env.reset()
        if self.output == 'numpy':
            torch.cuda.synchronize()
        return self._out(obs)

    def step_async(self, actions):
        custom = self.runtime._cfg.reward_id == 0
        if self.output == 'numpy' and not custom:
            # numpy in / numpy out through the C-ABI host entry points: the actions are staged and H2D + kernel + ONE D2H
            # into a pinned block are enqueued now; step_wait only waits (the caller may do other work in between)
            a = actions if isinstance(actions, np.ndarray) and actions.dtype == np.float32 else np.asarray(actions, dtype=np.float32)
            self.runtime.engine.step_host_packed_begin(a)
            self._pending = ('host', None)
        else:
            self._pending = ('device', self.env.step(actions))                # enqueued on the current stream
        self.waiting = True

    def step_wait(self):
        kind, payload = self._pending
        self._pending, self.waiting = None, False
        names = self.runtime.task.reset_positions
        if kind == 'host':
            obs, rew, done, rid, t_idx, t_cause, t_obs = self.runtime.engine.step_host_packed_end()
            return obs, rew, done, LazyInfos(names, rid, done, t_idx, t_cause, t_obs)
        obs, rew, done, info = payload
        if self.output == 'torch':
            return obs, rew, done, info
        torch.cuda.synchronize()
        done_h = done.cpu().numpy()
        rid = info['reset_orientation'].cpu().numpy()
        term = info['terminal_observation'].cpu().numpy() if done_h.any() else None
        cause = info['cause'].cpu().numpy()
        return obs.cpu().numpy(), rew.cpu().numpy(), done_h, LazyInfos.from_dense(names, rid, done_h, term, cause)

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def get_state_info(self, states, actions):
        rew, done = self.runtime.task.get_state_info(states, actions)
        return (self._out(rew) if torch.is_tensor(rew) else np.asarray(rew),
                self._out(done) if torch.is_tensor(done) else np.asarray(done))

    def seed(self, seed=None):
        return self.env.seed(seed)

    def render(self, mode='human'):
        return self.env.render(mode)

    def close(self):
        if not self.closed:
            self.env.close()
            self.closed = True

    def _get_indices(self, indices):
        """subproc_vec_env.py:177-189 / vec_env.py:225-245: None = all envs, an int = that env, else an iterable."""
        if indices is None:
            return list(range(self.num_envs))
        if isinstance(indices, (int, np.integer)):
            indices = [int(indices)]
        out = [int(i) for i in indices]
        for i in out:
            if not -self.num_envs <= i < self.num_envs:
                raise IndexError(f'env index {i} out of range for {self.num_envs} envs')
        return out

    def get_attr(self, attr_name, indices=None):
        """One entry per selected env, as the reference returns (subproc_vec_env.py:150-156). All envs of the batch
        share ONE task / runtime object, so the entries are the same object."""
        value = getattr(self.env, attr_name)
        return [value for _ in self._get_indices(indices)]

    def set_attr(self, attr_name, value, indices=None):
        """The batch shares one runtime: an attribute cannot differ between envs, so only "all envs" is accepted."""
        if len(self._get_indices(indices)) != self.num_envs:
            raise ValueError('the envs of a CUDA batch share one runtime object: set_attr applies to all of them '
                             '(indices=None)')
        setattr(self.runtime, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        """Calls the method ONCE on the shared runtime and returns the result once per selected env
        (subproc_vec_env.py:166-175 returns one result per env)."""
        idx = self._get_indices(indices)
        result = getattr(self.env, method_name)(*args, **kwargs)
        return [result for _ in idx]

    @property
    def unwrapped(self):
        return self.runtime
