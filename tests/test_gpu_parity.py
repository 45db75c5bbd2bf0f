"""GPU: the CUDA env step (through the C ABI) against the oracle and the golden traces of the reference."""
import numpy as np
import pytest
import torch

from oracle import c_oracle

from helpers import (assert_close, assert_exact, golden_names, load_golden, obs_variant_names, oracle_kwargs,
                     pending_movement_names, pending_obs_names, pending_sequential_names, population_kwargs,
                     population_names, utility_names)

pytestmark = pytest.mark.gpu

# fp64 taps vs oracle: north_star asks 1e-5 relative; the kernels are fp64 and are held to 1e-9 (+ abs floor)
RTOL, ATOL = 1e-9, 1e-9
# production outputs are float32: rounding of the stored value only
RTOL32, ATOL32 = 2e-6, 1e-6
# the observation entry 'dr' (normalised SNR, variants.py:276-284) is evaluated in fp32 on the device: north_star
# tolerance for SINR is 1e-5 relative; held to 2e-6 here
RTOL_DR = 2e-6


def split_dr(obs, n_ue, n_bs):
    """(everything but 'dr', 'dr') of one env's packed observation"""
    obs = np.asarray(obs, dtype=np.float64)
    if obs.ndim == 1:       # central: connected[N*M] | dr[N*M] | utility[N]
        nm = n_ue * n_bs
        return np.concatenate([obs[:nm], obs[2 * nm:]]), obs[nm:2 * nm]
    return np.concatenate([obs[:, :n_bs], obs[:, 2 * n_bs:]], axis=1), obs[:, n_bs:2 * n_bs]


def make_env(cfg_kwargs, num_envs=1, seeds=None, **extra):
    from deepcomp_b200 import BatchedMobileEnv
    kw = dict(cfg_kwargs)
    seed = kw.pop('seed')
    kw.update(extra)
    if seeds is None:
        seeds = [seed]
    return BatchedMobileEnv(num_envs=num_envs, seeds=seeds, **kw)


def compare_step(env, dbg, want, k, what, step=True, num_ue=None):
    """dbg: debug dict of the CUDA env (all K envs); want: oracle trace of env k.  num_ue: UEs present (variable
    population): the state slabs are compared on those slots only, every output on all max_ues rows (padding = 0)."""
    st = env.get_state()
    n = want['mask'].shape[0] if num_ue is None else num_ue
    assert_exact(st['pos'][k][:n], want['pos'][:n], f'{what}.pos')
    assert_exact(env.mask_matrix(st['mask'])[k][:n], want['mask'][:n], f'{what}.mask')
    assert_exact(st['movement'][k][:n], want['movement'][:n], f'{what}.movement')
    assert_close(st['ewma'][k][:n], want['ewma'][:n], f'{what}.ewma', RTOL, ATOL)
    assert_close(dbg['dbg_snr'][k].cpu().numpy(), want['snr'], f'{what}.snr', RTOL, 0)
    assert_close(dbg['dbg_link_rate'][k].cpu().numpy(), want['link_rates'], f'{what}.link_rates', RTOL, ATOL)
    assert_close(dbg['dbg_curr_dr'][k].cpu().numpy(), want['curr_dr'], f'{what}.curr_dr', RTOL, ATOL)
    assert_close(dbg['dbg_utility'][k].cpu().numpy(), want['utility'], f'{what}.utility', RTOL, ATOL)
    n_ue, n_bs = want['mask'].shape
    if env.obs_variant is not None:
        # data-rate observation classes: every entry is fp64 on the device (tap) and rounded once (production output)
        assert_close(dbg['dbg_obs'][k].cpu().numpy(), want['obs'], f'{what}.obs64', RTOL, ATOL)
        assert_close(dbg['obs'][k].cpu().numpy(), want['obs'], f'{what}.obs32', RTOL32, ATOL32)
    else:
        # MaxNormEnv's 'dr' (variants.py:308-332) crosses zero at the connection threshold: absolute floor there
        dr_atol = 1e-9 if env.obs_norm == 'max' else 1e-30
        w_rest, w_dr = split_dr(want['obs'], n_ue, n_bs)
        g_rest, g_dr = split_dr(dbg['dbg_obs'][k].cpu().numpy(), n_ue, n_bs)
        assert_close(g_rest, w_rest, f'{what}.obs64', RTOL, ATOL)
        assert_close(g_dr, w_dr, f'{what}.obs64.dr', RTOL_DR, dr_atol)
        g_rest, g_dr = split_dr(dbg['obs'][k].cpu().numpy(), n_ue, n_bs)
        assert_close(g_rest, w_rest, f'{what}.obs32', RTOL32, ATOL32)
        assert_close(g_dr, w_dr, f'{what}.obs32.dr', RTOL_DR, dr_atol)
    if step:
        assert_exact(dbg['lost_conn'][k].cpu().numpy().astype(np.int32), want['lost_conn'], f'{what}.lost_conn')
        assert st['time'][k] == want['time']
        assert_close(dbg['dbg_reward'][k].cpu().numpy(), want['reward'], f'{what}.reward64', RTOL, ATOL)
        assert_close(dbg['reward'][k].cpu().numpy(), want['reward'], f'{what}.reward32', RTOL32, ATOL32)
        assert_close(dbg['dbg_sum_utility'][k].cpu().numpy(), want['sum_utility'], f'{what}.sum_utility', RTOL, ATOL)


@pytest.mark.parametrize('wide', [False, True], ids=['fused', 'wide'])
@pytest.mark.parametrize('name', golden_names() + utility_names() + obs_variant_names() + pending_movement_names())
def test_cuda_step_matches_reference_golden(name, wide, monkeypatch):
    """K=1, one launch per step, every recorded array of the reference trace; through the fused kernel (dcb_step.cu)
    and through the one-CTA-per-env kernel for large envs (dcb_wide.cu, forced here for the small golden shapes)."""
    if wide:
        monkeypatch.setenv('DCB_FORCE_WIDE', '1')
    cfg, z = load_golden(name)
    env = make_env(oracle_kwargs(cfg))
    t = 0
    for ep in range(cfg['episodes']):
        dbg = env.reset(debug=True)
        want = {k: z['reset_' + k][ep] for k in ('pos', 'mask', 'movement', 'ewma', 'snr', 'link_rates', 'curr_dr',
                                                 'utility', 'obs')}
        compare_step(env, dbg, want, 0, f'{name}.reset[{ep}]', step=False)
        for _ in range(cfg['steps']):
            a = torch.as_tensor(z['actions'][t][None, :].astype(np.int32), device='cuda')
            dbg = env.step(a, debug=True)
            want = {k: z['step_' + k][t] for k in ('pos', 'mask', 'movement', 'ewma', 'snr', 'link_rates', 'curr_dr',
                                                   'utility', 'obs', 'lost_conn', 'time', 'reward', 'sum_utility')}
            compare_step(env, dbg, want, 0, f'{name}.step[{t}]')
            t += 1
    env.check_errors()


@pytest.mark.parametrize('wide', [False, True], ids=['fused', 'wide'])
@pytest.mark.parametrize('name', population_names())
def test_cuda_variable_population_matches_reference_golden(name, wide, monkeypatch):
    """ue_arrival / new_ue_interval (base.py:429-443, 592-617) on envs with max_ues > num_ue: both episodes of the
    reference's trace, every recorded array, through both kernels.  The reset in between re-seeds the UEs of the list as
    it stands by list position and then restores the original list (base.py:132-143, 169-189): originals that moved up
    come back with another seed, originals that left continue their streams -- the second episode differs from the first."""
    if wide:
        monkeypatch.setenv('DCB_FORCE_WIDE', '1')
    cfg, z = load_golden(name)
    env = make_env(dict(oracle_kwargs(cfg), **population_kwargs(cfg)))
    step_keys = ('pos', 'mask', 'movement', 'ewma', 'snr', 'link_rates', 'curr_dr', 'utility', 'obs', 'lost_conn',
                 'time', 'reward', 'sum_utility')
    t = 0
    for ep in range(cfg['episodes']):
        dbg = env.reset(debug=True)
        want = {k: z['reset_' + k][ep] for k in ('pos', 'mask', 'movement', 'ewma', 'snr', 'link_rates', 'curr_dr',
                                                 'utility', 'obs')}
        assert env.active_ues == int(z['reset_num_ue'][ep])
        compare_step(env, dbg, want, 0, f'{name}.reset[{ep}]', step=False, num_ue=env.active_ues)
        for _ in range(cfg['steps']):
            a = torch.as_tensor(z['actions'][t][None, :].astype(np.int32), device='cuda')
            dbg = env.step(a, debug=True)
            assert env.active_ues == int(z['step_num_ue'][t]), (name, t)
            compare_step(env, dbg, {k: z['step_' + k][t] for k in step_keys}, 0, f'{name}.step[{t}]',
                         num_ue=env.active_ues)
            t += 1
    env.check_errors()


@pytest.mark.parametrize('wide', [False, True], ids=['fused', 'wide'])
@pytest.mark.parametrize('name', pending_obs_names())
def test_cuda_datarate_observation_classes_match_reference_golden(name, wide, monkeypatch):
    """CentralNormDrEnv / CentralDrEnv (central.py:75-140, variants.py:42-250): the shared rate every UE gets or would get
    from every BS, all env_config options of the data-rate class (auto / numeric cut-off, required-rate subtraction, total
    rate, UEs per BS, distances now and after the next step), every recorded array of the reference's traces.  These
    fused kernel's observers run one step behind the physics warps, which keep that step's per-(env, BS) aggregates for
    them per step parity."""
    if wide:
        monkeypatch.setenv('DCB_FORCE_WIDE', '1')
    cfg, z = load_golden(name)
    env = make_env(oracle_kwargs(cfg))
    assert env.kernel_name == ('dcb_wide_kernel' if wide else 'dcb_step_kernel') and env.obs_size == z['step_obs'].shape[1]
    t = 0
    for ep in range(cfg['episodes']):
        dbg = env.reset(debug=True)
        want = {k: z['reset_' + k][ep] for k in ('pos', 'mask', 'movement', 'ewma', 'snr', 'link_rates', 'curr_dr',
                                                 'utility', 'obs')}
        compare_step(env, dbg, want, 0, f'{name}.reset[{ep}]', step=False)
        for _ in range(cfg['steps']):
            a = torch.as_tensor(z['actions'][t][None, :].astype(np.int32), device='cuda')
            dbg = env.step(a, debug=True)
            want = {k: z['step_' + k][t] for k in ('pos', 'mask', 'movement', 'ewma', 'snr', 'link_rates', 'curr_dr',
                                                   'utility', 'obs', 'lost_conn', 'time', 'reward', 'sum_utility')}
            compare_step(env, dbg, want, 0, f'{name}.step[{t}]')
            t += 1
    env.check_errors()


@pytest.mark.parametrize('variant,opts', [('normdr', None), ('datarate', dict(dr_cutoff='auto', sub_req_dr=True,
                                          curr_dr_obs=True, ues_at_bs_obs=True, dist_obs=True, next_dist_obs=True))])
def test_datarate_observation_fragments_match_the_python_oracle(variant, opts):
    """the data-rate observation classes on a K-env batch in multi-step fragments (the pipelined path of the fused kernel:
    observers one step behind, aggregates per step parity), every step of every env against the Python oracle"""
    from oracle.deepcomp_oracle import OracleEnv
    from deepcomp_b200 import BatchedMobileEnv, env_seeds
    K, N, M, T = 9, 20, 7, 24
    W, H, bs = c_oracle_grid(M)
    seeds = env_seeds(3, K, N)
    kw = dict(kind='central', n_ue=N, bs_xy=bs, map_wh=(W, H), sharing=['resource-fair', 'rate-fair', 'proportional-fair',
              'max-cap', 'rate-fair', 'proportional-fair', 'max-cap'], velocities='slow', reward='avg', episode_length=T,
              obs_variant=variant, obs_opts=opts)
    env = BatchedMobileEnv(num_envs=K, seeds=seeds, **kw)
    assert env.kernel_name == 'dcb_step_kernel'
    orcs = [OracleEnv(seed=int(sd), **kw) for sd in seeds]
    obs0 = env.reset().cpu().numpy()
    for k, o in enumerate(orcs):
        assert_close(obs0[k], o.reset_trace()['obs'], f'reset.env{k}', RTOL32, ATOL32)
    a = np.random.default_rng(4).integers(0, M + 1, (T, K, N)).astype(np.int32)
    out = env.step_many(torch.as_tensor(a, device='cuda'))
    obs, rew = out['obs'].cpu().numpy(), out['reward'].cpu().numpy()
    for k, o in enumerate(orcs):
        for t in range(T):
            w = o.step(a[t, k])
            assert_close(obs[t, k], w['obs'], f'obs[{t}].env{k}', RTOL32, ATOL32)
            assert_close(rew[t, k], w['reward'], f'reward[{t}].env{k}', RTOL32, ATOL32)
    env.check_errors()


@pytest.mark.parametrize('wide', [False, True], ids=['fused', 'wide'])
@pytest.mark.parametrize('name', pending_sequential_names())
def test_cuda_sequential_env_matches_reference_golden(name, wide, monkeypatch):
    """SeqMultiAgentMobileEnv (multi_agent.py:110-179): one UE acts per call, nothing moves until the last UE of the round
    has acted (dcb_step_no_move), the call returns the NEXT UE's observation row and reward -- every recorded array of the
    reference's trace, two envs in lockstep, through both kernels."""
    if wide:
        monkeypatch.setenv('DCB_FORCE_WIDE', '1')
    cfg, z = load_golden(name)
    kw = oracle_kwargs(cfg)
    assert kw.pop('sequential')
    env = make_env(kw, num_envs=2, seeds=[cfg['seed'], cfg['seed']])
    n_ue, n_bs = cfg['n_ue'], len(cfg['bs_xy'])
    dbg = env.reset(debug=True)
    nxt = 0
    for k in range(2):
        want = {key: z['reset_' + key][0] for key in ('pos', 'mask', 'movement', 'ewma', 'snr', 'link_rates', 'curr_dr',
                                                      'utility')}
        want['obs'] = dbg['dbg_obs'][k].cpu().numpy()                  # the trace holds the current UE's row only ...
        compare_step(env, dbg, want, k, f'{name}.reset', step=False)
        assert_close(dbg['dbg_obs'][k, nxt].cpu().numpy(), z['reset_obs'][0], f'{name}.reset.obs', RTOL_DR, 1e-30)
    for t in range(cfg['steps']):
        a = torch.as_tensor(np.stack([z['actions'][t]] * 2).astype(np.int32), device='cuda')
        dbg = env.step_sequential(a, debug=True)
        nxt = dbg['ue_index']
        st = env.get_state()
        for k in range(2):
            what = f'{name}.step[{t}].env{k}'
            assert_exact(st['pos'][k], z['step_pos'][t], what + '.pos')
            assert_exact(env.mask_matrix(st['mask'])[k], z['step_mask'][t], what + '.mask')
            assert_exact(st['movement'][k], z['step_movement'][t], what + '.movement')
            assert st['time'][k] == z['step_time'][t]
            assert_exact(dbg['lost_conn'][k].cpu().numpy().astype(np.int32), z['step_lost_conn'][t], what + '.lost_conn')
            assert_close(st['ewma'][k], z['step_ewma'][t], what + '.ewma', RTOL, ATOL)
            assert_close(dbg['dbg_snr'][k].cpu().numpy(), z['step_snr'][t], what + '.snr', RTOL, 0)
            assert_close(dbg['dbg_link_rate'][k].cpu().numpy(), z['step_link_rates'][t], what + '.link_rates', RTOL, ATOL)
            assert_close(dbg['dbg_curr_dr'][k].cpu().numpy(), z['step_curr_dr'][t], what + '.curr_dr', RTOL, ATOL)
            assert_close(dbg['dbg_utility'][k].cpu().numpy(), z['step_utility'][t], what + '.utility', RTOL, ATOL)
            assert_close(dbg['dbg_sum_utility'][k].cpu().numpy(), z['step_sum_utility'][t], what + '.sum_utility', RTOL, ATOL)
            w_rest, w_dr = split_dr(z['step_obs'][t][None, :], 1, n_bs)
            g_rest, g_dr = split_dr(dbg['dbg_obs'][k, nxt].cpu().numpy()[None, :], 1, n_bs)
            assert_close(g_rest, w_rest, what + '.obs64', RTOL, ATOL)
            assert_close(g_dr, w_dr, what + '.obs64.dr', RTOL_DR, 1e-30)
            assert_close(dbg['obs'][k, nxt].cpu().numpy(), z['step_obs'][t], what + '.obs32', RTOL32, ATOL32)
            assert_close(dbg['dbg_reward'][k, nxt].cpu().numpy(), z['step_reward'][t], what + '.reward64', RTOL, ATOL)
            assert_close(dbg['reward'][k, nxt].cpu().numpy(), z['step_reward'][t], what + '.reward32', RTOL32, ATOL32)
    env.check_errors()


def test_variable_population_fragments_equal_single_steps():
    """step_many across arrivals / departures (one launch per stretch without an event) == single steps, K envs."""
    from deepcomp_b200 import BatchedMobileEnv
    W, H, bs = c_oracle_grid(5)
    arrival = {5: 2, 6: 1, 20: -2, 21: 3, 40: -4}
    kw = dict(num_envs=5, n_ue=3, max_ues=10, bs_xy=bs, map_wh=(W, H), kind='multi', seed=11, episode_length=60,
              ue_arrival=arrival, new_ue_interval=17)
    e1, e2 = BatchedMobileEnv(**kw), BatchedMobileEnv(**kw)
    a = torch.randint(0, 6, (60, 5, 10), dtype=torch.int32, device='cuda', generator=torch.Generator('cuda').manual_seed(1))
    e1.reset(); e2.reset()
    f = e1.step_many(a, info=True)
    for t in range(60):
        obs, rew, _, info = e2.step(a[t])
        assert torch.equal(obs, f['obs'][t]) and torch.equal(rew, f['reward'][t]), t
        assert torch.equal(info['lost_conn'], f['lost_conn'][t]) and torch.equal(info['sum_utility'], f['sum_utility'][t])
    want = 3 + sum(arrival.values()) + 3                       # + the interval arrivals at 17, 34, 51
    assert e1.active_ues == e2.active_ues == want
    s1, s2 = e1.get_state(), e2.get_state()
    n = e1.active_ues
    assert np.array_equal(s1['pos'][:, :n], s2['pos'][:, :n]) and np.array_equal(s1['mask'][:, :n], s2['mask'][:, :n])
    # envs differ in WHO left (per-env draws of the global `random` module) and where the newcomers appeared
    assert len({tuple(p.ravel()) for p in s1['pos'][:, :n]}) == 5


@pytest.mark.parametrize('kind', ['central', 'multi'])
@pytest.mark.parametrize('n_ue,n_bs,K', [(5, 3, 7), (50, 10, 33), (200, 20, 5), (1, 1, 3), (33, 64, 4)])
def test_cuda_batch_matches_c_oracle(kind, n_ue, n_bs, K):
    """K envs with distinct seeds, 40 steps, every step compared against the C restatement."""
    from deepcomp_b200 import env_seeds
    W, H, bs = c_oracle_grid(n_bs)
    seeds = env_seeds(1000, K, n_ue)
    kw = dict(kind=kind, n_ue=n_ue, bs_xy=bs, map_wh=(W, H), sharing='mixed', velocities='slow', reward='avg',
              episode_length=40)
    env = make_env(dict(kw, seed=0), num_envs=K, seeds=seeds)
    orcs = [c_oracle.COracleEnv(seed=int(s), **kw) for s in seeds]
    dbg = env.reset(debug=True)
    for k, o in enumerate(orcs):
        compare_step(env, dbg, o.reset_trace(), k, f'reset.env{k}', step=False)
    rng = np.random.default_rng(5)
    for t in range(40):
        a = rng.integers(0, n_bs + 1, (K, n_ue)).astype(np.int32)
        dbg = env.step(torch.as_tensor(a, device='cuda'), debug=True)
        for k, o in enumerate(orcs):
            compare_step(env, dbg, o.step(a[k]), k, f'step[{t}].env{k}')
    env.check_errors()


@pytest.mark.parametrize('kind', ['central', 'multi'])
@pytest.mark.parametrize('n_ue,n_bs,K,steps,force', [(1000, 50, 2, 12, False), (600, 33, 2, 12, False),
                                                     (50, 10, 9, 40, True), (5, 3, 7, 40, True), (33, 64, 3, 40, True),
                                                     (1, 1, 3, 20, True)])
def test_wide_kernel_matches_c_oracle(kind, n_ue, n_bs, K, steps, force, monkeypatch):
    """BASELINE config 4 shape (1000 UE x 50 BS) and other shapes through dcb_wide.cu, every step against the C oracle."""
    from deepcomp_b200 import env_seeds
    if force:
        monkeypatch.setenv('DCB_FORCE_WIDE', '1')
    W, H, bs = c_oracle_grid(n_bs)
    seeds = env_seeds(1000, K, n_ue)
    kw = dict(kind=kind, n_ue=n_ue, bs_xy=bs, map_wh=(W, H), sharing='mixed', velocities='slow', reward='avg',
              episode_length=steps)
    env = make_env(dict(kw, seed=0), num_envs=K, seeds=seeds)
    assert env.launch_geometry['grid'] == K              # one CTA per env
    orcs = [c_oracle.COracleEnv(seed=int(s), **kw) for s in seeds]
    dbg = env.reset(debug=True)
    for k, o in enumerate(orcs):
        compare_step(env, dbg, o.reset_trace(), k, f'reset.env{k}', step=False)
    rng = np.random.default_rng(11)
    for t in range(steps):
        a = rng.integers(0, n_bs + 1, (K, n_ue)).astype(np.int32)
        dbg = env.step(torch.as_tensor(a, device='cuda'), debug=True)
        for k, o in enumerate(orcs):
            compare_step(env, dbg, o.step(a[k]), k, f'step[{t}].env{k}')
    env.check_errors()


@pytest.mark.parametrize('kind', ['central', 'multi'])
@pytest.mark.parametrize('n_ue,n_bs,K,near', [(20, 10, 3, False), (12, 5, 2, True), (300, 40, 2, False)])
def test_interference_extension_matches_c_restatement(kind, n_ue, n_bs, K, near):
    """EXTENSION, not parity-graded against the reference (which is SNR only, station.py:122-127): SINR_b = snr_b / (1 +
    sum of the other base stations' snr) everywhere the SNR is used -- against its only oracle, the C restatement.
    `near`: some UEs sit on top of / within a metre of a base station, where the interference term is not ~1e-8 but
    dominates every other link of that UE."""
    from deepcomp_b200 import env_seeds
    W, H, bs = c_oracle_grid(n_bs)
    seeds = env_seeds(77, K, n_ue)
    init_pos = None
    velocities = 'slow'
    if near:
        init_pos = [(bs[0][0], bs[0][1]), (bs[1][0] + 0.25, bs[1][1]), (bs[2][0], bs[2][1] - 0.9)] + \
            [('random', 'random')] * (n_ue - 3)
        velocities = [0, 0, 0] + ['slow'] * (n_ue - 3)
    kw = dict(kind=kind, n_ue=n_ue, bs_xy=bs, map_wh=(W, H), sharing='mixed', velocities=velocities, reward='avg',
              episode_length=30, init_pos=init_pos)
    env = make_env(dict(kw, seed=0), num_envs=K, seeds=seeds, interference=True)
    assert env.kernel_name == 'dcb_wide_kernel'
    orcs = [c_oracle.COracleEnv(seed=int(s), interference=True, **kw) for s in seeds]
    plain = c_oracle.COracleEnv(seed=int(seeds[0]), **kw)
    dbg = env.reset(debug=True)
    for k, o in enumerate(orcs):
        compare_step(env, dbg, o.reset_trace(), k, f'reset.env{k}', step=False)
    plain.reset_trace()
    rng = np.random.default_rng(6)
    differs = 0.0
    for t in range(30):
        a = rng.integers(0, n_bs + 1, (K, n_ue)).astype(np.int32)
        dbg = env.step(torch.as_tensor(a, device='cuda'), debug=True)
        for k, o in enumerate(orcs):
            w = o.step(a[k])
            compare_step(env, dbg, w, k, f'step[{t}].env{k}')
            if k == 0:
                differs = max(differs, float(np.max(np.abs(w['snr'] / plain.step(a[0])['snr'] - 1.0))))
    assert differs > 1e-9          # the extension is not a no-op: the SINR differs from the SNR beyond the parity bar
    env.check_errors()


@pytest.mark.parametrize('kind,reward', [('central', 'avg'), ('multi', 'avg'), ('multi', 'min'), ('multi', 'sum')])
def test_interference_fragment_equals_single_steps(kind, reward):
    """Interference extension: T fused steps in one launch == T single-step launches (the per-step interference pass writes
    the 'dr' segment of each step's observation; the row phase the rest), with an on-device episode reset in between; an
    observe-only launch returns the observation the last step returned."""
    from deepcomp_b200 import BatchedMobileEnv, env_seeds
    n_ue, n_bs, K, T = 70, 36, 3, 20
    W, H, bs = c_oracle_grid(n_bs)
    kw = dict(num_envs=K, n_ue=n_ue, bs_xy=bs, map_wh=(W, H), kind=kind, reward=reward, seeds=env_seeds(11, K, n_ue),
              episode_length=8, auto_reset=True, interference=True)
    a = torch.randint(0, n_bs + 1, (T, K, n_ue), dtype=torch.int32, device='cuda',
                      generator=torch.Generator('cuda').manual_seed(5))
    e1, e2 = BatchedMobileEnv(**kw), BatchedMobileEnv(**kw)
    e1.reset(); e2.reset()
    f = e1.step_many(a)
    for t in range(T):
        obs, rew, _, info = e2.step(a[t])
        assert torch.equal(obs, f['obs'][t]) and torch.equal(rew, f['reward'][t]), t
        assert torch.equal(info['lost_conn'], f['lost_conn'][t]), t
    assert torch.equal(e1.observe(), f['obs'][T - 1]) and torch.equal(e2.observe(), f['obs'][T - 1])
    assert float(f['obs'].abs().sum()) > 0 and bool(torch.isfinite(f['obs']).all())
    s1, s2 = e1.get_state(), e2.get_state()
    for key in ('pos', 'mask', 'ewma', 'movement', 'time'):
        assert np.array_equal(s1[key], s2[key]), key
    e1.check_errors(); e2.check_errors()


def test_wide_fragment_equals_single_steps():
    """T fused steps in one launch of the wide kernel == T single-step launches (state carried in registers vs slabs),
    with an on-device episode reset in the middle."""
    from deepcomp_b200 import BatchedMobileEnv, env_seeds
    n_ue, n_bs, K, T = 520, 12, 3, 30
    W, H, bs = c_oracle_grid(n_bs)
    kw = dict(num_envs=K, n_ue=n_ue, bs_xy=bs, map_wh=(W, H), kind='multi', seeds=env_seeds(7, K, n_ue),
              episode_length=20, auto_reset=True)
    a = torch.randint(0, n_bs + 1, (T, K, n_ue), dtype=torch.int32, device='cuda',
                      generator=torch.Generator('cuda').manual_seed(3))
    e1, e2 = BatchedMobileEnv(**kw), BatchedMobileEnv(**kw)
    e1.reset(); e2.reset()
    f = e1.step_many(a)
    for t in range(T):
        obs, rew, _, info = e2.step(a[t])
        assert torch.equal(obs, f['obs'][t]) and torch.equal(rew, f['reward'][t])
        assert torch.equal(info['lost_conn'], f['lost_conn'][t])
    s1, s2 = e1.get_state(), e2.get_state()
    for key in ('pos', 'mask', 'ewma', 'movement', 'time'):
        assert np.array_equal(s1[key], s2[key]), key


def c_oracle_grid(n_bs):
    from oracle.deepcomp_oracle import grid_layout
    return grid_layout(n_bs)
