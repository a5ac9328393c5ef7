"""Thin object wrapper over the C-ABI handle (``os2r_env``): device buffers are torch tensors, the
work is done by libos2r.so. This is plumbing, not a fallback: every method raises if the CUDA
library or a CUDA device is missing."""
import ctypes as C
import sys

import numpy as np
import torch

from .. import _capi


class Engine:
    """N monopod environments resident on one GPU.

    ``step`` consumes / produces CUDA tensors on the current torch stream (asynchronous);
    ``step_host`` consumes / produces numpy arrays (H2D + kernel + D2H inside, synchronous).
    """

    def __init__(self, compiled_model, task_cfg: _capi.TaskCfg, n_envs: int, device: int = 0, seed: int = 0,
                 first_env_id: int = 0, precision: int = 32, tuning: dict = None):
        self.lib = _capi.load_library()
        if not torch.cuda.is_available():
            raise _capi.Os2rError('no CUDA device visible to torch; the monopod step path has no CPU fallback')
        self.compiled = compiled_model
        self.model = compiled_model.struct
        self.task_cfg = task_cfg
        self.n_envs = int(n_envs)
        self.device_index = int(device)
        self.device = torch.device('cuda', self.device_index)
        self.obs_dim = int(task_cfg.obs_dim)
        self.state_width = _capi.state_width(self.model)
        self.params_width = _capi.params_width(self.model)
        handle = C.c_void_p()
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)        # make sure the primary context exists
            tune = _capi.Tuning(**(tuning or {}))      # scheduling / verification knobs; never change a result
            _capi.check(self.lib.os2r_create_tuned(C.byref(self.model), C.byref(task_cfg), self.n_envs,
                                                   int(first_env_id), self.device_index, int(seed) & (2 ** 64 - 1),
                                                   int(precision), C.byref(tune), C.byref(handle)), self.lib)
        self.handle = handle
        N, D = self.n_envs, self.obs_dim
        kw = dict(device=self.device)
        self.obs = torch.zeros((N, D), dtype=torch.float32, **kw)
        self.reward = torch.zeros(N, dtype=torch.float32, **kw)
        self.done_u8 = torch.zeros(N, dtype=torch.uint8, **kw)
        self.terminal_obs = torch.zeros((N, D), dtype=torch.float32, **kw)
        self.info = torch.zeros((N, 2), dtype=torch.int32, **kw)
        self._host_pool = {}
        self._packed_layouts = {}
        self._packed_inflight = None
        self._action_buffer = None

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, 'handle', None) is not None and self.handle:
            self.lib.os2r_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def seed(self, seed: int):
        _capi.check(self.lib.os2r_seed(self.handle, int(seed) & (2 ** 64 - 1)), self.lib)

    # ------------------------------------------------------------------ hot path
    def reset(self, mask: torch.Tensor = None) -> torch.Tensor:
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            assert m.shape == (self.n_envs,)
        _capi.check(self.lib.os2r_reset(self.handle, C.c_void_p(m.data_ptr()) if m is not None else None,
                                        C.c_void_p(self.obs.data_ptr()), self._stream()), self.lib)
        return self.obs

    def step(self, actions: torch.Tensor, want_terminal_obs: bool = True, want_info: bool = True):
        """actions: float32 CUDA tensor [N, 2]. Returns views of the persistent output tensors."""
        if actions.dtype != torch.float32 or not actions.is_cuda or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(actions.shape) != (self.n_envs, 2):
            raise ValueError(f'actions must have shape ({self.n_envs}, 2), got {tuple(actions.shape)}')
        _capi.check(self.lib.os2r_step(
            self.handle, C.c_void_p(actions.data_ptr()), C.c_void_p(self.obs.data_ptr()),
            C.c_void_p(self.reward.data_ptr()), C.c_void_p(self.done_u8.data_ptr()),
            C.c_void_p(self.terminal_obs.data_ptr()) if want_terminal_obs else None,
            C.c_void_p(self.info.data_ptr()) if want_info else None, self._stream()), self.lib)
        return self.obs, self.reward, self.done_u8, self.info

    def _host_array(self, key, shape, dtype):
        """A page-locked numpy array for a step_host output. Arrays are recycled ONLY when the caller holds no
        reference to them any more (refcount check), so results are never overwritten behind the user's back,
        yet the steady state allocates nothing (fresh 2 MB arrays per step cost ~1 ms in page faults) and the
        GPU DMAs straight into the array the user receives (no staging copy)."""
        pool = self._host_pool.setdefault(key, [])
        for k in range(len(pool)):
            if sys.getrefcount(pool[k]) <= 2:      # the pool's reference + getrefcount's argument
                return pool[k]
        t = torch.empty(shape, dtype=dtype).pin_memory()
        arr = t.numpy()                            # keeps `t` alive through .base
        if len(pool) < 64:
            pool.append(arr)
        return arr

    @property
    def action_buffer(self) -> np.ndarray:
        """The handle's page-locked ``[N, 2]`` float32 action array (``os2r_host_action_buffer``). A caller that writes
        its actions here and passes THIS array to ``step_host`` / ``step_host_packed*`` skips the staging memcpy: the
        host-to-device copy reads it directly. Do not write to it between ``step_host_packed_begin`` and ``_end``."""
        if self._action_buffer is None:
            ptr = C.POINTER(C.c_float)()
            _capi.check(self.lib.os2r_host_action_buffer(self.handle, C.byref(ptr)), self.lib)
            self._action_buffer = np.ctypeslib.as_array(ptr, shape=(self.n_envs, 2))
        return self._action_buffer

    def step_host(self, actions: np.ndarray, want_terminal_obs: bool = False, want_info: bool = False):
        """numpy in / numpy out: H2D + kernel + D2H inside os2r_step_host. Returned arrays belong to the caller."""
        a = np.ascontiguousarray(actions, dtype=np.float32)
        if a.shape != (self.n_envs, 2):
            raise ValueError(f'actions must have shape ({self.n_envs}, 2), got {a.shape}')
        N, D = self.n_envs, self.obs_dim
        obs = self._host_array('obs', (N, D), torch.float32)
        rew = self._host_array('rew', (N,), torch.float32)
        done = self._host_array('done', (N,), torch.bool)
        term = self._host_array('term', (N, D), torch.float32) if want_terminal_obs else None
        info = self._host_array('info', (N, 2), torch.int32) if want_info else None
        p = lambda x: x.ctypes.data_as(C.c_void_p) if x is not None else None
        _capi.check(self.lib.os2r_step_host(self.handle, p(a), p(obs), p(rew), p(done), p(term), p(info)), self.lib)
        return obs, rew, done, term, info

    def step_host_packed_begin(self, actions: np.ndarray, prefix_records: int = None):
        """First half of ``step_host_packed`` (``os2r_step_host_packed_begin``): stages the actions and enqueues
        H2D + kernel + D2H, then returns without waiting — ``VecEnv.step_async``. ``actions`` may be reused at once."""
        a = np.ascontiguousarray(actions, dtype=np.float32)
        if a.shape != (self.n_envs, 2):
            raise ValueError(f'actions must have shape ({self.n_envs}, 2), got {a.shape}')
        N = self.n_envs
        if prefix_records is None:
            prefix_records = min(N, max(256, N // 64))
        L = self._packed_layouts.get(prefix_records)
        if L is None:
            L = _capi.PackedLayout()
            _capi.check(self.lib.os2r_packed_layout_get(self.handle, int(prefix_records), C.byref(L)), self.lib)
            self._packed_layouts[prefix_records] = L
        block = self._host_array(('block', prefix_records), (int(L.total_bytes),), torch.uint8)
        _capi.check(self.lib.os2r_step_host_packed_begin(self.handle, a.ctypes.data_as(C.c_void_p),
                                                         block.ctypes.data_as(C.c_void_p), int(prefix_records)), self.lib)
        self._packed_inflight = (block, L, int(prefix_records))

    def step_host_packed_end(self):
        """Second half (``os2r_step_host_packed_end``): waits for the step enqueued by ``step_host_packed_begin`` and
        returns ``(obs[N,D], reward[N], done[N] bool, reset_id[N] uint8, term_idx[k], term_cause[k], term_obs[k,D])``,
        views of ONE page-locked block (recycled only once the caller has dropped every view); the ``term_*`` arrays
        describe the k envs that finished an episode in this step."""
        if self._packed_inflight is None:
            raise _capi.Os2rError('step_host_packed_end without a step in flight')
        block, L, prefix_records = self._packed_inflight
        self._packed_inflight = None
        n_term = C.c_int32(0)
        _capi.check(self.lib.os2r_step_host_packed_end(self.handle, C.byref(n_term)), self.lib)
        N, D = self.n_envs, self.obs_dim
        obs = block[L.obs:L.obs + N * D * 4].view(np.float32).reshape(N, D)
        rew = block[L.reward:L.reward + N * 4].view(np.float32)
        done = block[L.done:L.done + N].view(np.bool_)
        rid = block[L.reset_id:L.reset_id + N]
        k, rw = int(n_term.value), int(L.record_words)
        if k <= prefix_records:
            rec = block[L.term_records:L.term_records + k * rw * 4].view(np.int32).reshape(k, rw)
        else:   # rare (e.g. every env hits the TimeLimit in the same step): fetch the full list
            rec = np.empty((k, rw), dtype=np.int32)
            _capi.check(self.lib.os2r_fetch_terminal_records(self.handle, 0, k, rec.ctypes.data_as(C.c_void_p)), self.lib)
        return obs, rew, done, rid, rec[:, 0], rec[:, 1], rec[:, 2:].view(np.float32)

    def step_host_packed(self, actions: np.ndarray, prefix_records: int = None):
        """numpy in / numpy out at full batch size: ONE device-to-host copy into one page-locked block."""
        self.step_host_packed_begin(actions, prefix_records)
        return self.step_host_packed_end()

    # ------------------------------------------------------------------ state access
    def get_state(self) -> np.ndarray:
        out = np.empty((self.n_envs, self.state_width), dtype=np.float64)
        _capi.check(self.lib.os2r_get_state(self.handle, out.ctypes.data_as(C.c_void_p)), self.lib)
        return out

    def set_state(self, state: np.ndarray):
        s = np.ascontiguousarray(state, dtype=np.float64)
        assert s.shape == (self.n_envs, self.state_width), s.shape
        _capi.check(self.lib.os2r_set_state(self.handle, s.ctypes.data_as(C.c_void_p)), self.lib)

    def get_params(self) -> np.ndarray:
        out = np.empty((self.n_envs, self.params_width), dtype=np.float64)
        _capi.check(self.lib.os2r_get_params(self.handle, out.ctypes.data_as(C.c_void_p)), self.lib)
        return out

    def set_params(self, params: np.ndarray):
        p = np.ascontiguousarray(params, dtype=np.float64)
        assert p.shape == (self.n_envs, self.params_width), p.shape
        _capi.check(self.lib.os2r_set_params(self.handle, p.ctypes.data_as(C.c_void_p)), self.lib)

    def get_episode(self):
        steps = np.empty(self.n_envs, dtype=np.int32)
        ret = np.empty(self.n_envs, dtype=np.float64)
        _capi.check(self.lib.os2r_get_episode(self.handle, steps.ctypes.data_as(C.c_void_p),
                                              ret.ctypes.data_as(C.c_void_p), None, None), self.lib)
        return steps, ret

    def get_reset_ids(self) -> np.ndarray:
        ids = np.empty(self.n_envs, dtype=np.int32)
        _capi.check(self.lib.os2r_get_episode(self.handle, None, None, ids.ctypes.data_as(C.c_void_p), None), self.lib)
        return ids

    def get_episode_counters(self) -> np.ndarray:
        """Per-env episode counter = counter word of the env's RNG stream (checkpoint / resume)."""
        ep = np.empty(self.n_envs, dtype=np.uint32)
        _capi.check(self.lib.os2r_get_episode(self.handle, None, None, None, ep.ctypes.data_as(C.c_void_p)), self.lib)
        return ep

    def set_episode(self, steps=None, returns=None, reset_ids=None, episodes=None):
        """``os2r_set_episode``: restore the per-env bookkeeping (any argument may be None = unchanged)."""
        def arr(x, dt):
            if x is None:
                return None, None
            a = np.ascontiguousarray(x, dtype=dt)
            assert a.shape == (self.n_envs,), a.shape
            return a, a.ctypes.data_as(C.c_void_p)
        keep = [arr(steps, np.int32), arr(returns, np.float64), arr(reset_ids, np.int32), arr(episodes, np.uint32)]
        _capi.check(self.lib.os2r_set_episode(self.handle, *[p for _, p in keep]), self.lib)

    def set_randomization(self, task_cfg: _capi.TaskCfg):
        """``os2r_set_randomization``: new ranges / switches for the draws made at the next resets."""
        _capi.check(self.lib.os2r_set_randomization(self.handle, C.byref(task_cfg)), self.lib)

    def set_stats(self, stats: dict):
        s = _capi.Stats(**{k: stats[k] for k, _ in _capi.Stats._fields_})
        _capi.check(self.lib.os2r_stats_write(self.handle, C.byref(s)), self.lib)

    def stats(self, clear: bool = False) -> dict:
        s = _capi.Stats()
        _capi.check(self.lib.os2r_stats_read(self.handle, C.byref(s), int(clear)), self.lib)
        return {name: getattr(s, name) for name, _ in _capi.Stats._fields_}

    def kernel_info(self) -> dict:
        vals = [C.c_int32() for _ in range(6)]
        _capi.check(self.lib.os2r_kernel_info(self.handle, *[C.byref(v) for v in vals]), self.lib)
        return dict(zip(('block_threads', 'grid_blocks', 'regs_per_thread', 'local_bytes_per_thread',
                         'resident_blocks_per_sm', 'envs_per_thread'), (v.value for v in vals)))

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.os2r_kernel_launches(self.handle))


def measure_fp32_peak(device: int = 0):
    lib = _capi.load_library()
    tf, mhz = C.c_double(), C.c_double()
    _capi.check(lib.os2r_measure_fp32_peak(int(device), C.byref(tf), C.byref(mhz)), lib)
    return tf.value, mhz.value
