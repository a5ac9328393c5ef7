"""GPU tests (-m gpu) of the reference-facing API: pytest translations of the reference's script-style
tests/tests_general.py:12-160 (check_registered_envs, test_random_rollout, single_process,
single_process_fixed_hip, test_monopod_model) plus the vector entry point and the torch policy loop."""
import functools

import numpy as np
import pytest
import torch

import gym_os2r                                   # the drop-in alias package
from gym_os2r import randomizers
from gym_os2r.common import make_env_from_id, make_mp_envs
from gym_os2r.rewards import RewardBase

pytestmark = pytest.mark.gpu
ENV_IDS = [s.id for s in gym_os2r._impl._gymshim.registry.all() if 'Monopod' in s.id]


@pytest.mark.parametrize('env_id', ENV_IDS)
def test_registered_env_contract(env_id):
    """tests_general.py:12-58 — spaces, dtypes, scalar reward, bool done, render before/after close."""
    env = randomizers.monopod.MonopodEnvRandomizer(env=functools.partial(make_env_from_id, env_id=env_id))
    env.seed(42)
    ob = env.reset()
    assert env.observation_space.contains(ob), ob
    assert ob.dtype == env.observation_space.dtype == np.float64
    a = env.action_space.sample()
    observation, reward, done, info = env.step(a)
    assert env.observation_space.contains(observation)
    assert np.isscalar(reward) and isinstance(done, bool) and observation.dtype == np.float64
    assert info['reset_orientation'] in env.unwrapped.task.reset_positions
    for mode in env.metadata.get('render.modes', []):
        env.render(mode=mode)
    env.close()
    for mode in env.metadata.get('render.modes', []):
        env.render(mode=mode)


@pytest.mark.parametrize('env_id', ENV_IDS)
def test_random_rollout(env_id):
    """tests_general.py:61-78"""
    env = randomizers.monopod.MonopodEnvRandomizer(env=functools.partial(make_env_from_id, env_id=env_id))
    env.seed(42)
    ob = env.reset()
    for _ in range(10):
        assert env.observation_space.contains(ob)
        a = env.action_space.sample()
        assert env.action_space.contains(a)
        ob, _reward, done, _info = env.step(a)
        if done:
            break
    env.close()


@pytest.mark.parametrize('task_mode,dim', [('free_hip', 10), ('fixed_hip', 8)])
def test_single_process(task_mode, dim):
    """tests_general.py:81-125"""
    make_env = functools.partial(make_env_from_id, env_id='Monopod-balance-v1', task_mode=task_mode)
    env = randomizers.monopod_no_rand.MonopodEnvNoRandomizer(env=make_env)
    env.seed(42)
    observation = env.reset()
    assert len(observation) == dim
    assert env.get_state_info(observation, [0, 0])[1] is False
    action = env.action_space.sample()
    after, reward, done, _ = env.step(action)
    assert env.get_state_info(after, action)[0] == reward
    moving = [i for n, i in env.unwrapped.task.observation_index.items() if 'yaw' not in n and 'boom_connector' not in n]
    assert all(after[moving] != observation[moving])
    # Task methods are views of the fused launch (reference: recomputed three times per step)
    task = env.unwrapped.task
    assert np.array_equal(task.get_observation(), after) and task.get_reward() == reward and task.is_done() == done
    env.close()


def test_monopod_model_reset_determinism():
    """tests_general.py:128-160"""
    def resets(randomizer, poses, k):
        env = randomizer(env=functools.partial(make_env_from_id, env_id='Monopod-balance-v1', reset_positions=poses))
        env.seed(42)
        out = np.vstack([env.reset() for _ in range(k)])
        env.close()
        return out
    assert not (np.diff(resets(randomizers.monopod_no_rand.MonopodEnvNoRandomizer, ['stand', 'ground'], 9), axis=0) == 0).all()
    assert (np.diff(resets(randomizers.monopod_no_rand.MonopodEnvNoRandomizer, ['stand'], 7), axis=0) == 0).all()
    assert not (np.diff(resets(randomizers.monopod.MonopodEnvRandomizer, ['stand'], 2), axis=0) == 0).all()


def test_custom_reward_class_runs_on_device():
    """examples/minimal_example.py:15-29 — a user RewardBase subclass (torch-evaluated on the GPU)."""
    class HeightV0(RewardBase):
        def __init__(self, observation_index, normalized):
            super().__init__(observation_index, normalized)
            self.supported_task_modes = ['fixed_hip']

        def calculate_reward(self, obs, actions):
            return obs[..., self.observation_index['planarizer_pitch_joint_pos']] * 2 + 1
    env = randomizers.monopod_no_rand.MonopodEnvNoRandomizer(env=functools.partial(
        make_env_from_id, env_id='Monopod-balance-v1', task_mode='fixed_hip', reward_class=HeightV0))
    env.reset()
    ob, reward, done, _ = env.step([0.1, -0.1])
    assert reward == pytest.approx(ob[2] * 2 + 1, abs=1e-6)
    env.close()
    # batched + auto-reset: the reward is evaluated on the pre-reset observation of every env, and the episode
    # statistics accumulate the user's reward (the kernel itself saw reward 0)
    N, limit = 256, 5
    envs = make_mp_envs('Monopod-balance-v1', N, 3, randomizers.monopod_no_rand.MonopodEnvNoRandomizer,
                        task_mode='fixed_hip', reward_class=HeightV0, max_episode_steps=limit)
    envs.output = 'torch'
    envs.reset()
    total = torch.zeros(N, dtype=torch.float64, device='cuda')
    for t in range(limit):
        obs, rew, done, info = envs.step(torch.zeros((N, 2), device='cuda'))
        term = info['terminal_observation']
        assert torch.allclose(rew, (term[:, 2] * 2 + 1), atol=1e-6)
        total += rew.double()
    assert done.all()
    st = envs.runtime.stats()
    assert st['episodes'] == N and st['sum_return'] == pytest.approx(float(total.sum()), rel=1e-9)
    envs.close()


def test_vec_env_numpy_and_torch_paths_agree():
    """make_mp_envs (common/__init__.py:27-53): numpy in/out, auto-reset with terminal_observation, seeds by rank."""
    N = 512
    kw = dict(task_mode='fixed_hip', max_episode_steps=6)
    e_np = make_mp_envs('Monopod-balance-v1', N, 3, randomizers.monopod.MonopodEnvRandomizer, **kw)
    e_t = make_mp_envs('Monopod-balance-v1', N, 3, randomizers.monopod.MonopodEnvRandomizer, **kw)
    e_t.output = 'torch'
    o1 = e_np.reset()
    o2 = e_t.reset()
    assert isinstance(o1, np.ndarray) and o1.shape == (N, 8) and np.array_equal(o1, o2.cpu().numpy())
    assert e_np.observation_space.shape == (8,) and e_np.num_envs == N
    rng = np.random.RandomState(0)
    for t in range(7):
        a = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        obs, rew, done, infos = e_np.step(a)
        obs_t, rew_t, done_t, info_t = e_t.step(torch.as_tensor(a, device='cuda'))
        assert np.array_equal(obs, obs_t.cpu().numpy()) and np.array_equal(rew, rew_t.cpu().numpy())
        assert np.array_equal(done, done_t.cpu().numpy())
    assert done.dtype == np.bool_ and rew.shape == (N,)
    e_np.step_async(a); obs, rew, done, infos = e_np.step_wait()
    e_t.step(torch.as_tensor(a, device='cuda'))
    # step 6 hit the TimeLimit for every env: terminal observation differs from the returned reset observation
    e2 = make_mp_envs('Monopod-balance-v1', 4, 3, randomizers.monopod.MonopodEnvRandomizer, **kw)
    e2.reset()
    for t in range(6):
        obs, rew, done, infos = e2.step(np.zeros((4, 2), np.float32))
    assert done.all() and len(infos) == 4
    assert infos[0]['TimeLimit.truncated'] and infos[0]['terminal_observation'].shape == (8,)
    assert not np.array_equal(infos[0]['terminal_observation'], obs[0]) and infos[0]['reset_orientation'] == 'stand'
    r, d = e2.get_state_info(obs, np.zeros((4, 2)))
    assert r.shape == (4,) and d.shape == (4,) and not d.any()
    # get_attr / env_method return one entry per selected env (subproc_vec_env.py:150-175) and honour `indices`
    assert len(e2.get_attr('action_space')) == 4 and len(e2.get_attr('action_space', indices=[0, 2])) == 2
    assert len(e2.get_attr('action_space', indices=1)) == 1
    assert e2.env_method('get_state_info', obs[0], [0.0, 0.0], indices=[3])[0][1] is False
    with pytest.raises(IndexError):
        e2.get_attr('action_space', indices=[4])
    with pytest.raises(ValueError):
        e2.set_attr('foo', 1, indices=[0])
    e2.set_attr('foo', 1)
    assert e2.runtime.foo == 1
    for e in (e_np, e_t, e2):
        e.close()


def test_policy_rollout_stays_on_device_and_checkpoints():
    """BASELINE config 5 shape (small): torch MLP -> step -> obs, all CUDA tensors; get_state/set_state resume."""
    N = 4096
    envs = make_mp_envs('Monopod-hop-v1', N, 1, randomizers.monopod.MonopodEnvRandomizer)
    envs.output = 'torch'
    obs = envs.reset()
    policy = torch.nn.Sequential(torch.nn.Linear(10, 32), torch.nn.Tanh(), torch.nn.Linear(32, 2), torch.nn.Tanh()).cuda()
    with torch.no_grad():
        for _ in range(20):
            obs, rew, done, info = envs.step(policy(obs))
    assert obs.is_cuda and obs.shape == (N, 10) and rew.is_cuda and done.dtype == torch.bool
    assert torch.isfinite(obs).all() and obs.abs().max() <= 1.0
    rt = envs.runtime
    snap = rt.get_state()
    assert {'state', 'params', 'steps', 'returns', 'reset_ids', 'episodes', 'seed', 'stats'} <= set(snap)
    with torch.no_grad():
        a = policy(obs)
        o1 = envs.step(a)[0].clone()
        rt.set_state(snap)
        o2 = envs.step(a)[0].clone()
    assert torch.equal(o1, o2)          # resets included: the episode counters (RNG streams) were restored too
    # ... and into a NEW runtime (fresh engine): the rollout continues bit-identically
    envs2 = make_mp_envs('Monopod-hop-v1', N, 1, randomizers.monopod.MonopodEnvRandomizer)
    envs2.output = 'torch'
    envs2.runtime.set_state(snap)
    assert torch.equal(envs2.step(a)[0], o1)
    with torch.no_grad():
        for _ in range(5):
            a = policy(obs)
            obs = envs.step(a)[0]
            assert torch.equal(envs2.step(a)[0], obs)
    assert envs2.runtime.stats()['episodes'] == rt.stats()['episodes']
    envs2.close()
    # ScenarIO-style pokes (examples/ignition_interaction.py)
    model = rt.task.model
    assert len(model.joint_positions(['hip_joint', 'knee_joint'])) == 2
    assert rt.world.to_gazebo().set_gravity((0, 0, -5.0)) and rt.world.gravity()[2] == -5.0
    envs.close()


def test_joint_level_shims_and_single_iteration_runtime():
    """examples/ignition_interaction.py: per-joint reset / read-back through the ScenarIO-style shims, and a runtime whose
    env step is ONE physics iteration (agent_rate == physics_rate): ten such steps with a held action equal one ordinary
    env step (the reference's zero-order-hold loop, runtimes/gazebo_runtime.py:65-97)."""
    mk = lambda **kw: randomizers.monopod_no_rand.MonopodEnvNoRandomizer(env=functools.partial(
        make_env_from_id, env_id='Monopod-balance-v1', task_mode='fixed_hip', **kw))
    fine, coarse = mk(agent_rate=10000, physics_rate=10000), mk()
    assert fine.unwrapped.num_of_steps_per_run == 1 and coarse.unwrapped.num_of_steps_per_run == 10
    for env in (fine, coarse):
        env.reset()
        m = env.unwrapped.task.model
        assert m.get_joint('planarizer_pitch_joint').to_gazebo().reset_position(0.3)
        assert m.get_joint('hip_joint').to_gazebo().reset_position(-0.4)
        assert m.get_joint('hip_joint').joint_position()[0] == pytest.approx(-0.4, abs=1e-7)
    with pytest.raises(KeyError):
        fine.unwrapped.task.model.get_joint('boom_connector_joint')      # fixed in this model
    a = [0.3, -0.2]
    for _ in range(10):
        fine.step(a)
    coarse.step(a)
    names = ['planarizer_pitch_joint', 'hip_joint', 'knee_joint', 'planarizer_yaw_joint']
    qf, qc = fine.unwrapped.task.model.joint_positions(names), coarse.unwrapped.task.model.joint_positions(names)
    vf, vc = fine.unwrapped.task.model.joint_velocities(names), coarse.unwrapped.task.model.joint_velocities(names)
    np.testing.assert_array_equal(qf, qc)
    np.testing.assert_array_equal(vf, vc)
    assert fine.unwrapped.task.model.links_in_contact() == []
    fine.close(); coarse.close()


def test_handles_on_two_devices_in_one_process():
    """One process may hold handles on several GPUs: the > 48 KB dynamic shared memory opt-in of the wide-block step kernel
    is a per-device attribute and is made per handle (os2r_create under its device guard), so the SECOND device's launches
    work too; results are independent of the device."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (run with gpurun --gpus 2)')
    from gym_os2r_b200.runtimes.engine import Engine
    from helpers import make_config
    task, cm, cfg = make_config('free_hip', reward='HoppingV1', reset_randomized=True, randomize_params=True, pgs_tol=1e-6)
    N = 148 * 4 * 64 + 224                          # wide blocks: 5 DoF + 4 proxies need ~110 KB of shared memory per block
    out = []
    for dev in (0, 1):
        eng = Engine(cm, cfg, N, device=dev, seed=3)
        assert eng.kernel_info()['block_threads'] == 224
        eng.reset()
        g = torch.Generator(device=f'cuda:{dev}'); g.manual_seed(1)
        with torch.cuda.device(dev):
            for _ in range(3):
                obs, rew, done, _ = eng.step(torch.rand((N, 2), device=f'cuda:{dev}', generator=g) * 2 - 1)
            torch.cuda.synchronize(dev)
        out.append((eng, obs.cpu().numpy().copy(), eng.get_state()))
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    for eng, _, _ in out:
        eng.close()
