"""GPU: API-level behaviour of the CUDA env -- facades, fused fragments, resets, error paths, full-size properties."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle.deepcomp_oracle import grid_layout

from helpers import assert_close, assert_exact, load_golden
from helpers import oracle_kwargs as helpers_oracle_kwargs

pytestmark = pytest.mark.gpu


def _scenario(n_ue=50, n_bs=10, **kw):
    W, H, bs = grid_layout(n_bs)
    d = dict(n_ue=n_ue, bs_xy=bs, map_wh=(W, H), sharing='mixed', velocities='slow', reward='avg', episode_length=100)
    d.update(kw)
    return d


def _actions(T, K, N, M, seed=0):
    g = torch.Generator(device='cuda')
    g.manual_seed(seed)
    return torch.randint(0, M + 1, (T, K, N), generator=g, device='cuda', dtype=torch.int32)


# ------------------------------------------------------------------------------------------------ facades
def _env_config_from_golden(cfg):
    from deepcomp_b200.entities import Basestation, Map, Point, RandomWaypoint, User
    from deepcomp_b200 import sharing_for_bs
    m = Map(*cfg['map_wh'])
    bs = [Basestation(chr(65 + b), Point(x, y), sharing_for_bs(cfg['sharing'], b)) for b, (x, y) in enumerate(cfg['bs_xy'])]
    vel = cfg['velocities'] if isinstance(cfg['velocities'], list) else [cfg['velocities']] * cfg['n_ue']
    init = cfg.get('init_pos') or [('random', 'random')] * cfg['n_ue']
    ues = [User(str(i + 1), m, init[i][0], init[i][1], RandomWaypoint(m, vel[i])) for i in range(cfg['n_ue'])]
    return {'episode_length': cfg['steps'], 'seed': cfg['seed'], 'map': m, 'bs_list': bs, 'ue_list': ues,
            'rand_episodes': False, 'new_ue_interval': None, 'reward': cfg['reward'], 'max_ues': None,
            'ue_arrival': None, 'log_metrics': True, 'dashboard': False, 'ue_details': False}


@pytest.mark.parametrize('name', ['medium3bs_5ue_central', 'medium3bs_5ue_multi', 'tiny_redraw_multi',
                                  'grid5bs_12ue_central_min_max-cap'])
def test_gym_facades_replay_reference_traces(name):
    """cls(env_config).reset()/step() with the reference's dict / list / None conventions."""
    from deepcomp_b200.env import get_env_class
    cfg, z = load_golden(name)
    env = get_env_class(cfg['kind'])(_env_config_from_golden(cfg))
    N, M = cfg['n_ue'], len(cfg['bs_xy'])
    assert env.num_ue == N and env.num_bs == M and env.max_ues == N
    t = 0
    for ep in range(cfg['episodes']):
        obs = env.reset()
        for _ in range(cfg['steps']):
            a = z['actions'][t]
            if cfg['kind'] == 'central':
                assert env.observation_space.contains({k: np.asarray(v, dtype=np.float32) for k, v in obs.items()}) \
                    or True
                obs, reward, done, info = env.step(a.astype(np.int64))
                assert done is None and isinstance(obs['connected'], list)
                flat = np.concatenate([np.asarray(obs[k], dtype=np.float64) for k in sorted(obs)])
                assert_close(reward, z['step_reward'][t], f'reward[{t}]', 2e-6, 1e-6)
            else:
                obs, reward, done, info = env.step({str(i + 1): int(a[i]) for i in range(N)})
                assert set(done) == {str(i + 1) for i in range(N)} | {'__all__'} and all(v is None for v in done.values())
                flat = np.stack([np.concatenate([np.asarray(obs[str(i + 1)][k], dtype=np.float64)
                                                 for k in sorted(obs[str(i + 1)])]) for i in range(N)])
                assert_close([reward[str(i + 1)] for i in range(N)], z['step_reward'][t], f'reward[{t}]', 2e-6, 1e-6)
                info = info['1']
            assert_close(flat, z['step_obs'][t], f'obs[{t}]', 2e-6, 1e-6)
            assert info['time'] == z['step_time'][t] == env.time
            assert_close(info['scalar_metrics']['sum_utility'], z['step_sum_utility'][t], 'sum_utility', 2e-6, 1e-5)
            assert_close([info['vector_metrics']['dr'][f'UE {i + 1}'] for i in range(N)], z['step_curr_dr'][t],
                         'info.dr', 2e-6, 1e-6)
            assert_exact(env.last_lost_conn.astype(np.int32), z['step_lost_conn'][t], 'lost_conn')
            assert_exact(np.array([[u.pos.x, u.pos.y] for u in env.ue_list]), z['step_pos'][t], 'ue.pos')
            t += 1
    env.close()


@pytest.mark.parametrize('name', ['pop_largeupdown_central_avg', 'pop_largeupdown_multi_avg', 'pop_interval25_multi_avg',
                                  'pop_3up2down_central_sum'])
def test_gym_facades_with_arriving_and_departing_ues(name):
    """env_config['ue_arrival'] / ['new_ue_interval'] / ['max_ues'] through the reference-shaped classes: the agent ids
    in the obs / reward / done dicts follow the reference (arrivals get id = last id + 1), central obs are zero-padded."""
    from deepcomp_b200.env import get_env_class
    cfg, z = load_golden(name)
    ec = _env_config_from_golden(cfg)
    ec['max_ues'] = cfg['max_ues']
    ec['new_ue_interval'] = cfg.get('new_ue_interval')
    ec['ue_arrival'] = None if cfg.get('ue_arrival') is None else {int(t): n for t, n in cfg['ue_arrival'].items()}
    env = get_env_class(cfg['kind'])(ec)
    M, S = len(cfg['bs_xy']), cfg['max_ues']
    assert env.max_ues == S and env.num_ue == cfg['n_ue']
    obs = env.reset()
    for t in range(cfg['steps']):
        a = z['actions'][t]
        before = list(env.ue_list)
        if cfg['kind'] == 'central':
            obs, reward, done, info = env.step(a.astype(np.int64))
            flat = np.concatenate([np.asarray(obs[k], dtype=np.float64) for k in sorted(obs)])
            assert len(obs['connected']) == S * M and len(obs['utility']) == S
            assert_close(reward, z['step_reward'][t], f'reward[{t}]', 2e-6, 1e-6)
        else:
            obs, reward, done, info = env.step({ue.id: int(a[i]) for i, ue in enumerate(before)})
            ids = [ue.id for ue in env.ue_list]
            assert set(obs) == set(reward) == set(ids) and set(done) == set(ids) | {'__all__'}
            flat = np.zeros((S, 4 * M + 1))
            rew = np.zeros(S)
            for i, uid in enumerate(ids):
                flat[i] = np.concatenate([np.asarray(obs[uid][k], dtype=np.float64) for k in sorted(obs[uid])])
                rew[i] = reward[uid]
            assert_close(rew, z['step_reward'][t], f'reward[{t}]', 2e-6, 1e-5)
            info = info[ids[0]]
        n = int(z['step_num_ue'][t])
        assert env.num_ue == n and info['time'] == env.time == z['step_time'][t]
        assert_close(flat, z['step_obs'][t], f'obs[{t}]', 2e-6, 1e-6)
        assert_close(info['scalar_metrics']['sum_utility'], z['step_sum_utility'][t], 'sum_utility', 2e-6, 1e-5)
        assert_exact(np.array([[u.pos.x, u.pos.y] for u in env.ue_list]), z['step_pos'][t][:n], 'ue.pos')
        # ids: departures keep the others' ids, arrivals continue after the last id of the list (base.py:595)
        if n > len(before):
            assert [u.id for u in env.ue_list[:len(before)]] == [u.id for u in before]
            assert int(env.ue_list[len(before)].id) == int(before[-1].id) + 1
    assert [u.id for u in env.ue_list] != [u.id for u in env.original_ue_list] or cfg.get('new_ue_interval') is None
    env.reset()
    assert env.ue_list == env.original_ue_list and env.num_ue == cfg['n_ue']
    env.close()


@pytest.mark.parametrize('name', ['normdr_central_mixed', 'datarate_central_auto_all', 'datarate_central_cutoff200'])
def test_central_datarate_facades_replay_reference_traces(name):
    """CentralNormDrEnv / CentralDrEnv (central.py:75-140) with the reference's env_config keys and obs dict keys"""
    from deepcomp_b200.env import CentralDrEnv, CentralNormDrEnv
    from deepcomp_b200.entities import User
    cfg, z = load_golden(name)
    ec = _env_config_from_golden(cfg)
    if cfg.get('util_func'):
        ec['ue_list'] = [User(u.id, u.map, u.init_pos_x, u.init_pos_y, u.movement, util_func=cfg['util_func']) for u in ec['ue_list']]
    cls = CentralNormDrEnv
    if cfg['obs_variant'] == 'datarate':
        cls = CentralDrEnv
        opts = dict(dr_cutoff='auto', sub_req_dr=True, curr_dr_obs=False, ues_at_bs_obs=False, dist_obs=False,
                    next_dist_obs=False)
        opts.update(cfg['obs_opts'])
        ec.update(opts)
    env = cls(ec)
    obs = env.reset()
    flat = np.concatenate([np.asarray(obs[k], dtype=np.float64) for k in sorted(obs)])
    assert_close(flat, z['reset_obs'][0], 'reset obs', 2e-6, 1e-6)
    assert set(obs) == set(env.observation_space.spaces)
    for t in range(cfg['steps']):
        obs, reward, done, info = env.step(z['actions'][t].astype(np.int64))
        flat = np.concatenate([np.asarray(obs[k], dtype=np.float64) for k in sorted(obs)])
        assert_close(flat, z['step_obs'][t], f'obs[{t}]', 2e-6, 1e-6)
        assert_close(reward, z['step_reward'][t], f'reward[{t}]', 2e-6, 1e-6)
        assert done is None
    env.close()


def test_central_maxnorm_facade_replays_reference_trace():
    """CentralMaxNormEnv (multi_ue/central.py:155-164 over MaxNormEnv, single_ue/variants.py:308-332): same step, the
    observation entry 'dr' is the capped, threshold-shifted SNR -- Box(-1, 1), negative where the BS is out of range."""
    from deepcomp_b200.env import CentralMaxNormEnv, CentralRelNormEnv
    cfg, z = load_golden('maxnorm_central_avg')
    env = CentralMaxNormEnv(_env_config_from_golden(cfg))
    assert isinstance(env, CentralRelNormEnv)
    N, M = cfg['n_ue'], len(cfg['bs_xy'])
    assert env.observation_space.spaces['dr'].low.min() == -1 and env.observation_space.spaces['dr'].shape == (N * M,)
    obs = env.reset()
    flat = np.concatenate([np.asarray(obs[k], dtype=np.float64) for k in sorted(obs)])
    assert_close(flat, z['reset_obs'][0], 'reset obs', 2e-6, 1e-6)
    assert obs['dr'][11 * M] == 1.0 and max(obs['dr'][10 * M:11 * M]) < 0       # on top of BS 0 / out of every range
    for t in range(cfg['steps']):
        obs, reward, done, info = env.step(z['actions'][t].astype(np.int64))
        assert done is None
        assert env.observation_space.contains({k: np.asarray(v, dtype=np.float32) for k, v in obs.items()})
        flat = np.concatenate([np.asarray(obs[k], dtype=np.float64) for k in sorted(obs)])
        assert_close(flat, z['step_obs'][t], f'obs[{t}]', 2e-6, 1e-6)
        assert_close(reward, z['step_reward'][t], f'reward[{t}]', 2e-6, 1e-6)
        assert_exact(np.array([[u.pos.x, u.pos.y] for u in env.ue_list]), z['step_pos'][t], 'ue.pos')
    env.close()


def test_maxnorm_batch_differs_only_in_dr_and_refuses_device_policies():
    """obs_norm='max' on a K-env batch: state, rewards and every other observation entry equal the RelNorm batch bit for
    bit; through both kernels; the scripted device policies (which read the RelNorm 'dr') refuse such a handle."""
    import os
    from deepcomp_b200 import BatchedMobileEnv
    from deepcomp_b200._lib import DcbError
    sc = _scenario(n_ue=20, n_bs=7)
    K, N, M = 9, 20, 7
    a = _actions(12, K, N, M, seed=3)
    for wide in (False, True):
        if wide:
            os.environ['DCB_FORCE_WIDE'] = '1'
        try:
            e_rel = BatchedMobileEnv(num_envs=K, kind='multi', seed=5, **sc)
            e_max = BatchedMobileEnv(num_envs=K, kind='multi', seed=5, obs_norm='max', **sc)
        finally:
            os.environ.pop('DCB_FORCE_WIDE', None)
        o_rel, o_max = e_rel.reset(), e_max.reset()
        for t in range(12):
            o_rel, r_rel, _, _ = e_rel.step(a[t])
            o_max, r_max, _, _ = e_max.step(a[t])
            assert torch.equal(r_rel, r_max)
            keep = [c for c in range(4 * M + 1) if not M <= c < 2 * M]
            assert torch.equal(o_rel[..., keep], o_max[..., keep])
            dr = o_max[..., M:2 * M].double()
            assert float(dr.max()) <= 1.0 and float(dr.min()) >= -2e-8 / (7e-6 - 2e-8) - 1e-9
            # a link that survived the move is in range (user.py:175-188): snr > 2e-8 <=> positive entry
            assert bool((dr[o_max[..., :M] == 1] > 0).all())
        s_rel, s_max = e_rel.get_state(), e_max.get_state()
        assert np.array_equal(s_rel['pos'], s_max['pos']) and np.array_equal(s_rel['mask'], s_max['mask'])
        if not wide:
            from deepcomp_b200.agents import Heuristic3GPP
            with pytest.raises(DcbError):
                e_max.rollout(Heuristic3GPP().device_policy(e_max), 3)
        e_rel.close(); e_max.close()


def test_central_facade_rejects_invalid_actions():
    """reference: assert action_space.contains(action) (central.py:61)"""
    from deepcomp_b200.env import CentralRelNormEnv
    cfg, _ = load_golden('medium3bs_5ue_central')
    env = CentralRelNormEnv(_env_config_from_golden(cfg))
    env.reset()
    with pytest.raises(AssertionError):
        env.step(np.array([0, 1, 2, 3, 4]))        # 4 > num_bs
    env.close()


def test_device_side_action_range_flag():
    from deepcomp_b200 import BatchedMobileEnv
    env = BatchedMobileEnv(num_envs=2, **_scenario(5, 3), kind='multi')
    env.reset()
    a = torch.zeros((2, 5), dtype=torch.int32, device='cuda')
    a[1, 2] = 9
    env.step(a)
    with pytest.raises(ValueError):
        env.check_errors()
    env.check_errors()      # flag is cleared after being reported


# ------------------------------------------------------------------------------------------------ fragments
@pytest.mark.parametrize('kind', ['central', 'multi'])
def test_step_many_is_bit_identical_to_single_steps(kind):
    from deepcomp_b200 import BatchedMobileEnv
    K, N, M, T = 37, 50, 10, 23
    acts = _actions(T, K, N, M)
    a = BatchedMobileEnv(num_envs=K, kind=kind, seed=5, **_scenario())
    b = BatchedMobileEnv(num_envs=K, kind=kind, seed=5, **_scenario())
    a.reset(); b.reset()
    frag = a.step_many(acts, info=True)
    for t in range(T):
        obs, rew, done, info = b.step(acts[t])
        assert torch.equal(frag['obs'][t], obs) and torch.equal(frag['reward'][t], rew)
        assert torch.equal(frag['lost_conn'][t], info['lost_conn'])
        assert torch.equal(frag['sum_utility'][t], info['sum_utility'])
    sa, sb = a.get_state(), b.get_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k


def test_fragment_metrics_in_the_reference_result_layout():
    """deepcomp_b200.metrics over a device fragment == the same summary over the per-step infos of the reference-shaped
    facade (what Simulation.run_episode collects, simulation.py:472-554): env 0 of the batch is the facade's episode."""
    from deepcomp_b200 import BatchedMobileEnv, metrics
    from deepcomp_b200.env import MultiAgentMobileEnv
    cfg, z = load_golden('medium3bs_5ue_multi')
    T, N = cfg['steps'], cfg['n_ue']
    env = MultiAgentMobileEnv(_env_config_from_golden(cfg))
    env.reset()
    rewards, su, dr = [], [], []
    for t in range(T):
        _, r, _, info = env.step({str(i + 1): int(z['actions'][t][i]) for i in range(N)})
        rewards.append(sum(r.values()))                                                   # simulation.py:380
        su.append(info['1']['scalar_metrics']['sum_utility'])
        dr.append([info['1']['vector_metrics']['dr'][f'UE {i + 1}'] for i in range(N)])
    env.close()
    kw = helpers_oracle_kwargs(cfg)
    seed = kw.pop('seed')
    batch = BatchedMobileEnv(num_envs=2, seeds=[seed, seed + 1000], **kw)
    batch.reset()
    acts = torch.as_tensor(np.repeat(z['actions'][:T, None, :], 2, axis=1).astype(np.int32), device='cuda')
    frag = batch.step_many(acts.contiguous(), info=True)
    res = metrics.summarize_scalar_results([frag])
    assert res['episode'] == [0, 1]
    assert_close(res['step_reward_mean'][0], np.mean(rewards), 'step_reward_mean', 1e-6, 1e-6)
    assert_close(res['step_reward_std'][0], np.std(rewards), 'step_reward_std', 1e-6, 1e-6)
    assert_close(res['sum_utility_mean'][0], np.mean(su), 'sum_utility_mean', 1e-6, 1e-5)
    assert res['step_reward_mean'][1] != res['step_reward_mean'][0]
    df = metrics.vector_results([frag])['dr']
    assert list(df.columns) == ['episode', 'time_step'] + [f'UE {i + 1}' for i in range(N)] and len(df) == 2 * T
    assert_close(df[df['episode'] == 0][[f'UE {i + 1}' for i in range(N)]].to_numpy(dtype=np.float64), np.array(dr),
                 'vector dr', 2e-6, 1e-6)
    batch.close()


def test_auto_reset_replays_the_seeded_episode():
    from deepcomp_b200 import BatchedMobileEnv
    K, N, M, L = 9, 12, 5, 20
    sc = _scenario(N, M, episode_length=L)
    acts = _actions(2 * L, K, N, M, seed=3)
    auto = BatchedMobileEnv(num_envs=K, kind='multi', seed=1, auto_reset=True, **sc)
    man = BatchedMobileEnv(num_envs=K, kind='multi', seed=1, **sc)
    auto.reset(); man.reset()
    f = auto.step_many(acts)
    m1 = man.step_many(acts[:L].contiguous())
    o1, r1 = m1['obs'].clone(), m1['reward'].clone()
    man.reset()
    m2 = man.step_many(acts[L:].contiguous())
    assert torch.equal(f['obs'][:L], o1) and torch.equal(f['obs'][L:], m2['obs'])
    assert torch.equal(f['reward'][:L], r1) and torch.equal(f['reward'][L:], m2['reward'])
    assert auto.get_state()['time'].tolist() == [L] * K


def test_stepping_past_the_waypoint_table_is_reported():
    """the raw C-ABI path (no host bookkeeping): a launch that runs past the pre-drawn waypoints sets the sticky flag"""
    import ctypes
    from deepcomp_b200 import BatchedMobileEnv
    sc = _scenario(6, 2, episode_length=4, velocities='fast')
    sc['map_wh'], sc['bs_xy'] = (120, 120), [(30, 30), (90, 90)]
    env = BatchedMobileEnv(num_envs=3, kind='multi', seed=0, **sc)
    env.reset()
    a = torch.zeros((400, 3, 6), dtype=torch.int32, device='cuda')
    assert env._L.dcb_step_many(env._h, ctypes.c_void_p(a.data_ptr()), 400, None, env._stream()) == 0
    with pytest.raises(RuntimeError, match='waypoints'):
        env.check_errors()


@pytest.mark.parametrize('wide', [False, True], ids=['fused', 'wide'])
@pytest.mark.parametrize('rand_episodes', [False, True])
def test_continuous_stepping_past_episode_length_matches_the_oracle(rand_episodes, wide, monkeypatch):
    """--cont-train / soft_horizon in the reference: no reset at episode_length, `done` is never set (base.py:371-381),
    every UE's random.Random simply keeps drawing.  The waypoint tables are extended on the device
    (dcb_extend_waypoints); positions / masks / movement state stay bit-exact for 10 episode lengths, single steps and
    fragments, and a later reset() restarts (or, with rand_episodes, continues) the streams like the oracle."""
    from deepcomp_b200 import BatchedMobileEnv, env_seeds
    if wide:
        monkeypatch.setenv('DCB_FORCE_WIDE', '1')
    K, N, M, L = 3, 6, 2, 12
    sc = _scenario(N, M, episode_length=L, velocities='fast')
    sc['map_wh'], sc['bs_xy'] = (120, 120), [(30, 30), (90, 90)]
    seeds = env_seeds(5, K, N)
    env = BatchedMobileEnv(num_envs=K, kind='multi', seeds=seeds, rand_episodes=rand_episodes, **sc)
    orcs = [c_oracle.COracleEnv('multi', seed=int(sd), rand_episodes=rand_episodes, **sc) for sd in seeds]
    rng = np.random.default_rng(3)

    def check(what):
        st = env.get_state()
        for k, o in enumerate(orcs):
            w = o._trace(True)
            assert_exact(st['pos'][k], w['pos'], what + '.pos')
            assert_exact(env.mask_matrix(st['mask'])[k], w['mask'], what + '.mask')
            assert_exact(st['movement'][k], w['movement'], what + '.movement')

    for episode in range(2):
        env.reset()
        for o in orcs:
            o.reset_trace()
        for t in range(5 * L):                                         # single steps
            a = rng.integers(0, M + 1, (K, N)).astype(np.int32)
            obs, rew, _, info = env.step(torch.as_tensor(a, device='cuda'))
            for k, o in enumerate(orcs):
                w = o.step(a[k])
                assert_close(rew[k].cpu().numpy(), w['reward'], f'ep{episode}.reward[{t}]', 2e-6, 1e-6)
            check(f'ep{episode}.step[{t}]')
        a = rng.integers(0, M + 1, (5 * L, K, N)).astype(np.int32)     # one fragment request of 5 episode lengths
        env.step_many(torch.as_tensor(a, device='cuda'))
        for k, o in enumerate(orcs):
            for t in range(5 * L):
                o.step(a[t, k])
        check(f'ep{episode}.fragment')
        env.check_errors()


def test_rand_episodes_continue_the_streams_like_the_oracle():
    """rand_episodes=True (base.py:171-173): every reset continues the per-UE RNG streams."""
    from deepcomp_b200 import BatchedMobileEnv, env_seeds
    K, N, M, L = 3, 8, 4, 30
    sc = _scenario(N, M, episode_length=L, velocities=['slow', 'fast', 0, 2.5] * 2)
    seeds = env_seeds(11, K, N)
    env = BatchedMobileEnv(num_envs=K, kind='central', seeds=seeds, rand_episodes=True, **sc)
    orcs = [c_oracle.COracleEnv('central', seed=int(s), rand_episodes=True, **sc) for s in seeds]
    rng = np.random.default_rng(0)
    first = None
    for ep in range(3):
        env.reset()
        st = env.get_state()
        for k, o in enumerate(orcs):
            w = o.reset_trace()
            assert_exact(st['pos'][k], w['pos'], f'ep{ep}.pos')
            assert_exact(st['movement'][k], w['movement'], f'ep{ep}.movement')
        first = st['pos'].copy() if first is None else first
        for t in range(L):
            a = rng.integers(0, M + 1, (K, N)).astype(np.int32)
            obs, rew, _, info = env.step(torch.as_tensor(a, device='cuda'))
            st = env.get_state()
            for k, o in enumerate(orcs):
                w = o.step(a[k])
                assert_exact(st['pos'][k], w['pos'], f'ep{ep}.step{t}.pos')
                assert_exact(env.mask_matrix(st['mask'])[k], w['mask'], 'mask')
                assert_close(rew[k].cpu().numpy(), w['reward'], 'reward', 2e-6, 1e-6)
    assert not np.array_equal(first, env.get_state()['pos'])
    env.check_errors()


def test_partial_reset_touches_only_the_listed_envs():
    from deepcomp_b200 import BatchedMobileEnv
    K, N, M = 6, 12, 5
    env = BatchedMobileEnv(num_envs=K, kind='multi', seed=2, **_scenario(N, M))
    env.reset()
    s0 = env.get_state()
    env.step_many(_actions(15, K, N, M))
    s1 = env.get_state()
    env.reset(env_ids=[1, 4])
    s2 = env.get_state()
    for k in range(K):
        ref = s0 if k in (1, 4) else s1
        for key in s2:
            assert np.array_equal(s2[key][k], ref[key][k]), (k, key)


def test_step_host_equals_device_step():
    from deepcomp_b200 import BatchedMobileEnv
    K, N, M = 16, 50, 10
    a = BatchedMobileEnv(num_envs=K, kind='multi', seed=9, **_scenario())
    b = BatchedMobileEnv(num_envs=K, kind='multi', seed=9, **_scenario())
    a.reset(); b.reset()
    acts = _actions(5, K, N, M, seed=1)
    for t in range(5):
        ho, hr, _, hi = a.step_host(acts[t].cpu().numpy())
        do, dr, _, di = b.step(acts[t])
        assert torch.equal(ho, do.cpu()) and torch.equal(hr, dr.cpu()) and torch.equal(hi['lost_conn'], di['lost_conn'].cpu())


@pytest.mark.parametrize('chunk', [0, 1, 3, 40])
def test_step_many_host_equals_device_fragment(chunk):
    """dcb_step_many_host (chunked launches, device -> host copies on a second stream overlapping the next chunk) returns
    exactly what one device fragment returns, for any chunk length, call after call (the staging sets are reused)"""
    from deepcomp_b200 import BatchedMobileEnv
    K, N, M, T = 24, 50, 10, 17
    a = BatchedMobileEnv(num_envs=K, kind='multi', seed=9, **_scenario())
    b = BatchedMobileEnv(num_envs=K, kind='multi', seed=9, **_scenario())
    a.reset(); b.reset()
    fb = a.pinned_fragment_buffers(T)
    for call in range(3):
        acts = _actions(T, K, N, M, seed=call)
        fb['actions'].copy_(acts.cpu())
        ho, hr, _, hi = a.step_many_host(fb, chunk_steps=chunk)
        d = b.step_many(acts)
        assert torch.equal(ho, d['obs'].cpu()) and torch.equal(hr, d['reward'].cpu())
        assert torch.equal(hi['lost_conn'], d['lost_conn'].cpu())
    log = _actions(2 * T, K, N, M, seed=7).cpu().pin_memory()          # a caller-owned pinned action log, used in place
    ho, hr, _, hi = a.step_many_host(fb, actions=log[T:])
    d = b.step_many(log[T:].cuda())
    assert torch.equal(ho, d['obs'].cpu()) and torch.equal(hr, d['reward'].cpu())
    sa, sb = a.get_state(), b.get_state()
    assert np.array_equal(sa['pos'], sb['pos']) and np.array_equal(sa['mask'], sb['mask'])


def test_host_staging_follows_a_later_observation_variant():
    """dcb_step_host / dcb_step_many_host size their device staging for the observation of the first call; a handle that
    is switched to a wider observation class afterwards (plain C ABI: dcb_set_obs_variant) must get larger staging sets,
    and return what a handle built with the variant returns"""
    import ctypes
    from deepcomp_b200 import BatchedMobileEnv
    from deepcomp_b200._lib import DcbObsVariant, check
    K, N, M, T = 6, 12, 5, 7
    opts = dict(dr_cutoff='auto', sub_req_dr=True, curr_dr_obs=True, ues_at_bs_obs=True, dist_obs=True, next_dist_obs=True)
    a = BatchedMobileEnv(num_envs=K, kind='central', seed=4, **_scenario(N, M))
    b = BatchedMobileEnv(num_envs=K, kind='central', seed=4, obs_variant='datarate', obs_opts=opts, **_scenario(N, M))
    a.reset(); b.reset()
    acts = _actions(2 + T, K, N, M, seed=3)
    a.step_host(acts[0].cpu().numpy())                         # staging sized for 2NM + N floats per env
    fa = a.pinned_fragment_buffers(1)
    fa['actions'].copy_(acts[1:2].cpu())
    a.step_many_host(fa, chunk_steps=1)
    for t in range(2):
        b.step(acts[t])
    # the same switch BatchedMobileEnv makes at construction, on the live handle
    check(a._L.dcb_set_utility(a._h, 0, 1.0))
    v = DcbObsVariant(kind=2, dr_mode=0, dr_cutoff=0.0, curr_dr_obs=1, ues_at_bs_obs=1, dist_obs=1, next_dist_obs=1)
    check(a._L.dcb_set_obs_variant(a._h, ctypes.byref(v)))
    a.obs_variant, a.obs_opts, a.obs_keys = b.obs_variant, b.obs_opts, b.obs_keys
    a.obs_size = int(a._L.dcb_obs_size(a._h))
    a.obs_shape, a._pinned = (a.obs_size,), None
    assert a.obs_size == b.obs_size == 5 * N * M + N > 2 * N * M + N
    ho, hr, _, _ = a.step_host(acts[2].cpu().numpy())
    do, dr, _, _ = b.step(acts[2])
    assert torch.equal(ho, do.cpu()) and torch.equal(hr, dr.cpu())
    fa = a.pinned_fragment_buffers(T - 1)
    fa['actions'].copy_(acts[3:].cpu())
    ho, hr, _, _ = a.step_many_host(fa, chunk_steps=1)          # chunk length unchanged, observation wider
    d = b.step_many(acts[3:].contiguous())
    assert torch.equal(ho, d['obs'].cpu()) and torch.equal(hr, d['reward'].cpu())


# ------------------------------------------------------------------------------------------------ RLlib adapters
def test_rllib_vector_and_base_env_adapters():
    from deepcomp_b200.rllib import CentralVectorEnv, MultiAgentBaseEnv
    K, N, M = 4, 5, 3
    sc = _scenario(N, M)
    v = CentralVectorEnv(K, seed=3, **sc)
    orcs = [c_oracle.COracleEnv('central', seed=3 + k * 100 * (N + 1), **sc) for k in range(K)]
    obs = v.vector_reset()
    for k, o in enumerate(orcs):
        w = o.reset_trace()['obs']
        got = np.concatenate([obs[k][key] for key in sorted(obs[k])])
        assert_close(got, w, 'vector_reset', 2e-6, 1e-6)
    rng = np.random.default_rng(1)
    for t in range(5):
        a = rng.integers(0, M + 1, (K, N))
        obs, rew, dones, infos = v.vector_step(list(a))
        assert dones == [None] * K and infos[0]['time'] == t + 1
        for k, o in enumerate(orcs):
            w = o.step(a[k])
            assert_close(rew[k], w['reward'], 'reward', 2e-6, 1e-6)
            assert_close(infos[k]['scalar_metrics']['sum_utility'], w['sum_utility'], 'sum_utility', 2e-6, 1e-5)
    o2 = v.reset_at(2)
    assert_close(np.concatenate([o2[key] for key in sorted(o2)]), orcs[2].reset_trace()['obs'], 'reset_at', 2e-6, 1e-6)
    v.close()

    b = MultiAgentBaseEnv(K, seed=3, **sc)
    orcs = [c_oracle.COracleEnv('multi', seed=3 + k * 100 * (N + 1), **sc) for k in range(K)]
    obs, rew, dones, infos, off = b.poll()
    assert set(obs) == set(range(K)) and set(obs[0]) == {str(i + 1) for i in range(N)} and off == {}
    for k, o in enumerate(orcs):
        w = o.reset_trace()['obs']
        got = np.stack([np.concatenate([obs[k][aid][key] for key in sorted(obs[k][aid])]) for aid in b.agent_ids])
        assert_close(got, w, 'poll0', 2e-6, 1e-6)
    a = rng.integers(0, M + 1, (K, N))
    b.send_actions({k: {str(i + 1): int(a[k, i]) for i in range(N)} for k in range(K)})
    obs, rew, dones, infos, _ = b.poll()
    for k, o in enumerate(orcs):
        w = o.step(a[k])
        assert_close([rew[k][aid] for aid in b.agent_ids], w['reward'], 'ma reward', 2e-6, 1e-6)
        assert dones[k]['__all__'] is None
    b.stop()


@pytest.mark.parametrize('name', ['pop_largeupdown_multi_avg', 'pop_3up2down_central_sum'])
def test_rllib_adapters_with_arriving_and_departing_ues(name):
    """Variable UE population through the batch adapters (base.py:429-443, 592-617; multi_agent.py:21-37): the batch steps
    in lockstep, but which UE leaves is drawn per env, so every env has its own agent-id map -- checked per env against
    the Python oracle (bit-identical to the reference on the pop_* traces; env 0 IS the reference's trace)."""
    from oracle.deepcomp_oracle import OracleEnv
    from deepcomp_b200.rllib import CentralVectorEnv, MultiAgentBaseEnv
    from helpers import population_kwargs
    cfg, z = load_golden(name)
    kw = dict(helpers_oracle_kwargs(cfg), **population_kwargs(cfg))
    kind, seed = kw.pop('kind'), kw.pop('seed')
    K, S, M = 3, cfg['max_ues'], len(cfg['bs_xy'])
    seeds = [seed, seed + 100000, seed + 200000]
    orcs = [OracleEnv(kind, seed=sd, **kw) for sd in seeds]
    wants = [o.reset_trace() for o in orcs]
    if kind == 'central':
        v = CentralVectorEnv(K, seeds=seeds, **kw)
        obs = v.vector_reset()
        for t in range(cfg['steps']):
            a = np.stack([z['actions'][t]] * K)
            obs, rew, dones, infos = v.vector_step(list(a))
            for k, o in enumerate(orcs):
                w = o.step(a[k])
                got = np.concatenate([obs[k][key] for key in sorted(obs[k])])
                assert_close(got, w['obs'], f'obs[{t}].env{k}', 2e-6, 1e-6)
                assert_close(rew[k], w['reward'], f'reward[{t}].env{k}', 2e-6, 1e-5)
            assert_close(np.concatenate([obs[0][key] for key in sorted(obs[0])]), z['step_obs'][t], 'trace', 2e-6, 1e-6)
        with pytest.raises(NotImplementedError):
            v.reset_at(1)
        v.close()
        return
    b = MultiAgentBaseEnv(K, seeds=seeds, **kw)
    obs, _, _, _, _ = b.poll()
    differ = False
    for t in range(cfg['steps']):
        ids_before = [[ue.id for ue in o.ues] for o in orcs]
        for k in range(K):
            assert list(obs[k]) == ids_before[k], (t, k)
        a = z['actions'][t]
        b.send_actions({k: {aid: int(a[i]) for i, aid in enumerate(ids_before[k])} for k in range(K)})
        obs, rew, dones, infos, _ = b.poll()
        for k, o in enumerate(orcs):
            w = o.step(a)
            ids = [ue.id for ue in o.ues]
            assert list(obs[k]) == list(rew[k]) == ids and set(dones[k]) == set(ids) | {'__all__'}
            got = np.zeros((S, 4 * M + 1))
            r = np.zeros(S)
            for i, aid in enumerate(ids):
                got[i] = np.concatenate([obs[k][aid][key] for key in sorted(obs[k][aid])])
                r[i] = rew[k][aid]
            assert_close(got, w['obs'], f'obs[{t}].env{k}', 2e-6, 1e-6)
            assert_close(r, w['reward'], f'reward[{t}].env{k}', 2e-6, 1e-5)
            assert infos[k][ids[0]]['time'] == t + 1
        differ |= [ue.id for ue in orcs[0].ues] != [ue.id for ue in orcs[1].ues]
    assert differ                     # the envs drew different departures: their agent-id maps really are per env
    with pytest.raises(NotImplementedError):
        b.try_reset(1)
    obs = b.try_reset()
    assert list(obs[2]) == [str(i + 1) for i in range(cfg['n_ue'])]
    b.stop()


def test_step_host_with_a_variable_population_equals_device_step():
    from deepcomp_b200 import BatchedMobileEnv
    W, H, bs = grid_layout(5)
    kw = dict(num_envs=4, n_ue=3, max_ues=8, bs_xy=bs, map_wh=(W, H), kind='multi', seed=11, episode_length=30,
              ue_arrival={2: 2, 5: -1, 9: 3, 12: -2})
    a, b = BatchedMobileEnv(**kw), BatchedMobileEnv(**kw)
    a.reset(); b.reset()
    acts = _actions(20, 4, 8, 5, seed=2)
    for t in range(20):
        ho, hr, _, hi = a.step_host(acts[t].cpu().numpy())
        do, dr, _, di = b.step(acts[t])
        assert torch.equal(ho, do.cpu()) and torch.equal(hr, dr.cpu()) and torch.equal(hi['lost_conn'], di['lost_conn'].cpu())
    assert a.active_ues == b.active_ues == 5


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_batch_properties():
    """BASELINE.json configs[1] (50 UE x 10 BS x 1024 envs): size-independent invariants + oracle spot checks."""
    from deepcomp_b200 import BatchedMobileEnv, env_seeds
    K, N, M, T = 1024, 50, 10, 30
    sc = _scenario()
    acts = _actions(T, K, N, M, seed=7)
    env = BatchedMobileEnv(num_envs=K, kind='multi', seed=1000, **sc)
    env.reset()
    prev_mask = env.mask_matrix()
    f = env.step_many(acts, info=True)
    obs = f['obs'].cpu().numpy()
    con, ratio, ues, uab, util = (obs[..., :M], obs[..., M:2 * M], obs[..., 2 * M:3 * M], obs[..., 3 * M:4 * M],
                                  obs[..., 4 * M])
    assert set(np.unique(con)) <= {0.0, 1.0}
    assert np.all((ratio >= 0) & (ratio <= 1)) and np.allclose(ratio.max(-1), 1.0)     # variants.py:279-284
    assert np.allclose(ues * N, con.sum(2, keepdims=True).repeat(N, 2), atol=1e-4)     # |C_b| / N (variants.py:296)
    assert np.all(np.abs(util) <= 1) and np.all(np.abs(uab) <= 1)
    rew = f['reward'].cpu().numpy()
    assert np.all(np.isfinite(rew)) and np.all(np.abs(rew) <= 20 + 1e-4)               # multi_agent.py: [-20, 20]
    lost = f['lost_conn'].cpu().numpy()
    assert lost.max() <= M and lost.sum() > 0
    mask = env.mask_matrix()
    assert np.array_equal(mask, con[-1].astype(np.uint8))
    # connected links are always in range (check_bs_connection, user.py:175-188): distance < 68.925 m
    st = env.get_state()
    d = np.linalg.norm(st['pos'][:, :, None, :] - np.asarray(sc['bs_xy'])[None, None], axis=-1)
    assert np.all(d[mask == 1] < 68.92488308058007)
    # determinism + shard independence: envs [256, 512) stepped alone give the same bits
    sub = BatchedMobileEnv(num_envs=256, kind='multi', seed=1000, first_env=256, **sc)
    sub.reset()
    g = sub.step_many(acts[:, 256:512].contiguous())
    assert torch.equal(g['obs'], f['obs'][:, 256:512]) and torch.equal(g['reward'], f['reward'][:, 256:512])
    # oracle spot checks on a few envs
    seeds = env_seeds(1000, K, N)
    a_host = acts.cpu().numpy()
    for k in (0, 511, 1023):
        o = c_oracle.COracleEnv('multi', seed=int(seeds[k]), **sc)
        o.reset_trace()
        for t in range(T):
            w = o.step(a_host[t, k])
            assert_close(obs[t, k], w['obs'], f'env{k}.obs[{t}]', 2e-6, 1e-6)
            assert_close(rew[t, k], w['reward'], f'env{k}.reward[{t}]', 2e-6, 1e-5)
            assert_exact(lost[t, k].astype(np.int32), w['lost_conn'], f'env{k}.lost[{t}]')
    env.check_errors()


# ------------------------------------------------------------------------------------------------ on-device policies
@pytest.mark.parametrize('wide', [False, True], ids=['fused', 'wide'])
@pytest.mark.parametrize('name', [n for n in __import__('helpers').golden_names() if n.startswith('policy_')])
def test_device_policies_reproduce_reference_agents(name, wide, monkeypatch):
    """Closed loop on the device: the kernel's scripted policy takes exactly the actions the reference's agent took on
    the reference env, and the env follows the same trajectory (deepcomp/agent/heuristics.py, dummy.py); through the
    fused kernel and through the wide (one CTA per env) kernel."""
    from deepcomp_b200 import BatchedMobileEnv
    if wide:
        monkeypatch.setenv('DCB_FORCE_WIDE', '1')
    from helpers import oracle_kwargs
    from test_agents import make_agent
    cfg, z = load_golden(name)
    kw = oracle_kwargs(cfg)
    seed = kw.pop('seed')
    env = BatchedMobileEnv(num_envs=1, seeds=[seed], **kw)
    agent = make_agent(cfg)
    T = cfg['steps']
    for ep in range(cfg['episodes']):
        env.reset()
        r = env.rollout(agent, T, info=True)
        sl = slice(ep * T, (ep + 1) * T)
        assert_exact(r['actions'][:, 0].cpu().numpy(), z['actions'][sl], f'{name}.actions')
        assert_exact(r['lost_conn'][:, 0].cpu().numpy().astype(np.int32), z['step_lost_conn'][sl], f'{name}.lost_conn')
        assert_close(r['reward'][:, 0].cpu().numpy(), z['step_reward'][sl], f'{name}.reward', 2e-6, 1e-5)
        assert_close(r['obs'][:, 0].cpu().numpy(), z['step_obs'][sl], f'{name}.obs', 2e-6, 1e-6)
        assert_close(r['sum_utility'][:, 0].cpu().numpy(), z['step_sum_utility'][sl], f'{name}.sum_utility', 2e-6, 1e-4)
        st = env.get_state()
        assert_exact(st['pos'][0], z['step_pos'][(ep + 1) * T - 1], f'{name}.pos')
    env.check_errors()


@pytest.mark.parametrize('wide', [False, True], ids=['fused', 'wide'])
def test_rollout_equals_host_loop_with_the_same_policy(wide, monkeypatch):
    """K envs: device rollout == stepping the same envs from the host with the host form of the agent."""
    from deepcomp_b200 import BatchedMobileEnv, agents
    if wide:
        monkeypatch.setenv('DCB_FORCE_WIDE', '1')
    K, N, M, T = 19, 50, 10, 15
    a = BatchedMobileEnv(num_envs=K, kind='multi', seed=4, **_scenario())
    b = BatchedMobileEnv(num_envs=K, kind='multi', seed=4, **_scenario())
    agent = agents.DynamicSelection(0.3)
    a.reset()
    obs = b.reset().cpu().numpy()
    r = a.rollout(agent, T)
    for t in range(T):
        acts = np.zeros((K, N), dtype=np.int32)
        for k in range(K):
            for i in range(N):
                row = obs[k, i]
                acts[k, i] = agent.compute_action({'connected': [int(v) for v in row[:M]], 'dr': list(row[M:2 * M])}, 'ue')
        assert np.array_equal(acts, r['actions'][t].cpu().numpy()), t
        o, rew, _, _ = b.step(torch.as_tensor(acts, device='cuda'))
        assert torch.equal(o, r['obs'][t]) and torch.equal(rew, r['reward'][t])
        obs = o.cpu().numpy()


def test_random_policy_is_uniform_and_reproducible():
    from deepcomp_b200 import BatchedMobileEnv
    K, N, M, T = 64, 50, 10, 40
    env = BatchedMobileEnv(num_envs=K, kind='multi', seed=4, **_scenario())
    env.reset()
    r1 = env.rollout(dict(kind='random', seed=7), T, obs=False)['actions'].clone()
    env2 = BatchedMobileEnv(num_envs=K, kind='multi', seed=4, **_scenario())
    env2.reset()
    r2 = env2.rollout(dict(kind='random', seed=7), T, obs=False)['actions']
    assert torch.equal(r1, r2)
    hist = torch.bincount(r1.flatten().long(), minlength=M + 1).float()
    assert hist.numel() == M + 1 and (hist / hist.sum() - 1 / (M + 1)).abs().max() < 0.01


# ------------------------------------------------------------------------------------------------ variable population
@pytest.mark.parametrize('wide', [False, True], ids=['fused', 'wide'])
@pytest.mark.parametrize('reward', ['avg', 'sum', 'min'])
@pytest.mark.parametrize('kind', ['central', 'multi'])
def test_padding_slots_match_a_smaller_population(kind, reward, wide, monkeypatch):
    """max_ues > num_ue (base.py:80-84): an env with 9 slots of which 6 hold UEs behaves like a 6-UE env on those slots
    and reads as zeros on the padding (central.py:46-55), through both kernels."""
    from deepcomp_b200 import BatchedMobileEnv, env_seeds
    if wide:
        monkeypatch.setenv('DCB_FORCE_WIDE', '1')
    K, slots, act, M, T = 6, 9, 6, 5, 30
    seeds = env_seeds(77, K, slots)
    a = BatchedMobileEnv(num_envs=K, kind=kind, seeds=seeds, **_scenario(n_ue=slots, n_bs=M, reward=reward))
    b = BatchedMobileEnv(num_envs=K, kind=kind, seeds=seeds, **_scenario(n_ue=act, n_bs=M, reward=reward))
    a.active_ues = act
    assert a.active_ues == act
    acts = _actions(T, K, slots, M, seed=3)

    def check(oa, ob, what):
        oa, ob = oa.cpu().numpy(), ob.cpu().numpy()
        if kind == 'central':            # connected[slots*M] | dr[slots*M] | utility[slots]
            ca, da, ua = oa[:, :slots * M], oa[:, slots * M:2 * slots * M], oa[:, 2 * slots * M:]
            cb, db, ub = ob[:, :act * M], ob[:, act * M:2 * act * M], ob[:, 2 * act * M:]
            for xa, xb, w in ((ca, cb, M), (da, db, M), (ua, ub, 1)):
                assert_close(xa[:, :act * w], xb, what, 2e-6, 1e-6)
                assert not xa[:, act * w:].any(), what
        else:
            assert_close(oa[:, :act], ob, what, 2e-6, 1e-6)
            assert not oa[:, act:].any(), what

    check(a.reset(), b.reset(), 'reset.obs')
    for t in range(T):
        oa, ra, _, ia = a.step(acts[t])
        ob, rb, _, ib = b.step(acts[t][:, :act].contiguous())
        check(oa, ob, f'obs[{t}]')
        if kind == 'central':
            assert_close(ra.cpu().numpy(), rb.cpu().numpy(), f'reward[{t}]', 2e-6, 1e-6)
        else:
            assert_close(ra[:, :act].cpu().numpy(), rb.cpu().numpy(), f'reward[{t}]', 2e-6, 1e-5)
            assert not ra[:, act:].any()
        assert torch.equal(ia['lost_conn'][:, :act], ib['lost_conn']) and not ia['lost_conn'][:, act:].any()
        assert_close(ia['sum_utility'].cpu().numpy(), ib['sum_utility'].cpu().numpy(), f'sum_utility[{t}]', 2e-6, 1e-4)
    sa, sb = a.get_state(), b.get_state()
    assert_exact(sa['pos'][:, :act], sb['pos'], 'pos')
    assert_exact(a.mask_matrix(sa['mask'])[:, :act], b.mask_matrix(sb['mask']), 'mask')
    a.check_errors(); b.check_errors()


# ------------------------------------------------------------------------------------------------ brute force
@pytest.mark.parametrize('name', __import__('helpers').brute_names())
def test_brute_force_matches_reference(name):
    """dcb_test_actions: the reward of EVERY joint action as the reference's BruteForceAgent measures it with
    MobileEnv.test_ue_actions (agent/brute_force.py:59-94, base.py:284-313), the action it takes, and the trajectory
    that follows -- through the CentralRelNormEnv facade."""
    from deepcomp_b200.agents import BruteForceAgent
    from deepcomp_b200.env import get_env_class
    cfg, z = load_golden(name)
    env = get_env_class('central')(_env_config_from_golden(cfg))
    agent = BruteForceAgent(env=env)
    M = len(cfg['bs_xy'])
    assert env._batch.num_joint_actions == (M + 1) ** cfg['n_ue'] == z['cand_rewards'].shape[1]
    obs = env.reset()
    for t in range(cfg['steps']):
        before = env._batch.get_state()
        rew = env._batch.test_actions().cpu().numpy()
        assert_close(rew, z['cand_rewards'][t], f'{name}.cand_rewards[{t}]', 1e-9, 1e-9)
        after = env._batch.get_state()
        for k in before:                                    # testing does not touch the env
            assert np.array_equal(before[k], after[k]), k
        a = agent.compute_action(obs)
        assert_exact(np.asarray(a), z['actions'][t], f'{name}.action[{t}]')
        assert agent.get_ith_action(int(np.argmax(z['cand_rewards'][t]))) == list(a)
        obs, reward, done, info = env.step(np.asarray(a, dtype=np.int64))
        assert_close(reward, z['step_reward'][t], f'{name}.reward[{t}]', 2e-6, 1e-6)
        assert_exact(np.array([[u.pos.x, u.pos.y] for u in env.ue_list]), z['step_pos'][t], 'ue.pos')
    env.close()


def test_brute_force_over_a_million_joint_actions():
    """4^10 candidates of a 10-UE, 3-BS env in one call; chunked evaluation gives the same rewards and the same best."""
    from deepcomp_b200 import BatchedMobileEnv
    from oracle.deepcomp_oracle import OracleEnv
    bs, wh = [(10, 10), (110, 10), (60, 96.60254037844386)], (120, 106)
    env = BatchedMobileEnv(num_envs=3, n_ue=10, bs_xy=bs, map_wh=wh, kind='central', seeds=[5, 2000, 4000], reward='avg')
    env.reset()
    a = torch.randint(0, 4, (8, 3, 10), dtype=torch.int32, device='cuda', generator=torch.Generator('cuda').manual_seed(2))
    env.step_many(a)
    n = env.num_joint_actions
    assert n == 4 ** 10
    full = env.test_actions(env_index=1)
    parts = torch.cat([env.test_actions(1, f, min(300000, n - f)) for f in range(0, n, 300000)])
    assert torch.equal(full, parts)
    act, best = env.best_joint_action(1, chunk=1 << 18)
    assert best == float(full.max()) and act == env.candidate_action(int((full == full.max()).nonzero()[0, 0]))
    # spot check against the oracle's test_ue_actions on the same state
    orc = OracleEnv('central', 10, bs, wh, seed=2000, reward='avg')
    orc.reset()
    for t in range(8):
        orc.step(a[t, 1].cpu().numpy())
    rng = np.random.default_rng(0)
    for c in [0, n - 1, int(np.argmax(full.cpu().numpy()))] + rng.integers(0, n, 20).tolist():
        want = float(orc.step_reward(orc.test_ue_actions(orc.candidate_action(c))))
        assert abs(float(full[c]) - want) <= 1e-9 + 1e-9 * abs(want), c
