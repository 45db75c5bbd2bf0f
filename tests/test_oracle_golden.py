"""CPU: the oracle restatements (Python and C) against golden traces of the reference itself (tests/golden/)."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import deepcomp_oracle as po

from helpers import (GOLDEN_DIR, assert_close, assert_exact, brute_names, check_against_golden, golden_names, load_golden, obs_variant_names, oracle_kwargs, pending_movement_names, pending_obs_names,
                     pending_sequential_names,
                     population_kwargs, population_names, utility_names)


def test_anchor_known_answers():
    """Comment-level known answers of the reference (station.py:9,33-35; utility.py:49)."""
    z = np.load(f'{GOLDEN_DIR}/anchors.npz')
    snr = np.array([float(po.snr_of_distance(d)) for d in z['dist']])
    assert np.array_equal(snr, z['snr'])
    rate = np.array([float(po.BW * np.log2(1 + s)) for s in snr])
    assert np.array_equal(rate, z['rate_unshared'])
    util = np.array([float(po.log_utility(x)) for x in z['dr']])
    assert np.array_equal(util, z['log_utility'])
    assert po._CONST1 + po._CONST2 * np.log10(1 + po.EPSILON) == z['path_loss_1m']
    # "threshold corresponds roughly to a distance of 69m" (station.py:9)
    thr = po.threshold_distance()
    assert po.snr_of_distance(thr) > po.SNR_THRESHOLD >= po.snr_of_distance(np.nextafter(thr, 100))
    assert 68.92 < thr < 68.93
    # "1 Mbit range (46m)" (station.py:33)
    assert abs(po.BW * np.log2(1 + po.snr_of_distance(46)) - 1.0175) < 1e-3


def test_rng_known_answers():
    """MT19937 + CPython randint of the C restatement == random.Random draws recorded from the reference RNG."""
    z = np.load(f'{GOLDEN_DIR}/anchors.npz')
    ranges = [tuple(r) for r in z['rng_ranges'].tolist()]
    for i, seed in enumerate(z['rng_seeds'].tolist()):
        raw, ints = c_oracle.rng_draws(seed, 8, ranges)
        assert raw.tolist() == z['rng_raw'][i].tolist()
        assert ints.tolist() == z['rng_ints'][i].tolist()
    # env seed 42, UE "1" -> random.Random(142): start (298,249), v=3, first waypoint (259,94)
    raw, ints = c_oracle.rng_draws(142, 4, [(0, 300), (0, 300)])
    assert raw.tolist() == [2502198289, 2976738031, 2921394369, 2747885567] and ints.tolist() == [298, 249]
    _, ints = c_oracle.rng_draws(142, 0, [(1, 3), (10, 290), (10, 290)])
    assert ints.tolist() == [3, 259, 94]


PY_CASES = [n for n in golden_names() if not n.startswith('grid10bs_50ue')] + ['grid10bs_50ue_multi']


@pytest.mark.parametrize('name', PY_CASES)
def test_python_oracle_bit_identical_to_reference(name):
    cfg, z = load_golden(name)
    env = po.OracleEnv(**oracle_kwargs(cfg))
    # multi/'sum' sums over a Python set in hash order in the reference (user.py:238-244): O(1 ulp) freedom
    exact = not (cfg['kind'] == 'multi' and cfg['reward'] == 'sum')
    check_against_golden(env, cfg, z, exact_floats=exact)


@pytest.mark.parametrize('name', golden_names() + utility_names() + obs_variant_names())
def test_c_oracle_matches_reference(name):
    cfg, z = load_golden(name)
    env = c_oracle.COracleEnv(**oracle_kwargs(cfg))
    check_against_golden(env, cfg, z, exact_floats=False)


@pytest.mark.parametrize('name', population_names())
def test_python_oracle_variable_population_matches_reference(name):
    """ue_arrival / new_ue_interval on envs with max_ues > num_ue (base.py:433-443, 592-617; central.py:46-55), two
    episodes: the reset in between re-seeds by current list position and restores the original list (base.py:169-189)."""
    cfg, z = load_golden(name)
    env = po.OracleEnv(**oracle_kwargs(cfg), **population_kwargs(cfg))
    exact = not (cfg['kind'] == 'multi' and cfg['reward'] == 'sum')
    check_against_golden(env, cfg, z, exact_floats=exact)
    assert env.snapshot()['num_ue'] == int(z['step_num_ue'][-1])


@pytest.mark.parametrize('name', brute_names())
def test_python_oracle_brute_force_matches_reference(name):
    """BruteForceAgent (agent/brute_force.py:59-94) over MobileEnv.test_ue_actions (base.py:284-313): the reward of every
    joint action and the action taken, step by step."""
    cfg, z = load_golden(name)
    env = po.OracleEnv(**oracle_kwargs(cfg))
    env.reset()
    for t in range(cfg['steps']):
        rew = env.brute_force_rewards()
        assert_close(rew, z['cand_rewards'][t], f'{name}.cand_rewards[{t}]', 1e-12, 1e-12)
        a = env.candidate_action(int(np.argmax(rew)))
        assert_exact(np.asarray(a), z['actions'][t], f'{name}.action[{t}]')
        s = env.step(a)
        assert_exact(s['pos'], z['step_pos'][t], f'{name}.pos[{t}]')
        assert_exact(s['mask'], z['step_mask'][t], f'{name}.mask[{t}]')
        assert_close(s['reward'], z['step_reward'][t], f'{name}.reward[{t}]', 1e-12, 1e-12)


@pytest.mark.parametrize('name', utility_names())
def test_python_oracle_step_utility_matches_reference(name):
    """User.util_func = 'step' (CLI --util step; user.py:81-92, env/util/utility.py:23-33)"""
    cfg, z = load_golden(name)
    env = po.OracleEnv(**oracle_kwargs(cfg))
    exact = not (cfg['kind'] == 'multi' and cfg['reward'] == 'sum')
    check_against_golden(env, cfg, z, exact_floats=exact)


@pytest.mark.parametrize('name', obs_variant_names())
def test_python_oracle_maxnorm_observation_matches_reference(name):
    """MaxNormEnv.get_ue_obs (single_ue/variants.py:308-332): CentralMaxNormEnv (multi_ue/central.py:155-164) and the same
    composition over MultiAgentMobileEnv; incl. a UE on top of a BS (capped at 1) and one out of every BS's range."""
    cfg, z = load_golden(name)
    env = po.OracleEnv(**oracle_kwargs(cfg))
    check_against_golden(env, cfg, z, exact_floats=True)
    n, m = cfg['n_ue'], len(cfg['bs_xy'])
    dr = z['reset_obs'][0][n * m:2 * n * m].reshape(n, m) if cfg['kind'] == 'central' else z['reset_obs'][0][:, m:2 * m]
    assert dr[11, 0] == 1.0 and (dr[10] < 0).all()


@pytest.mark.parametrize('name', pending_obs_names())
def test_python_oracle_normdr_and_datarate_observations_match_reference(name):
    """CentralNormDrEnv (multi_ue/central.py:107-140, single_ue/variants.py:173-250) and CentralDrEnv (central.py:75-104,
    variants.py:42-170, all of its env_config options): the shared rate a UE gets or would get from every BS
    (station.py:204-220).  Oracle only -- the CUDA path does not offer these classes yet (DESIGN.md section 6e')."""
    cfg, z = load_golden(name)
    env = po.OracleEnv(**oracle_kwargs(cfg))
    check_against_golden(env, cfg, z, exact_floats=True)
    n, m = cfg['n_ue'], len(cfg['bs_xy'])
    width = z['step_obs'].shape[1]
    if cfg['obs_variant'] == 'normdr':
        assert width == 2 * n * m + n                       # connected | dr | dr_total
    elif cfg['obs_opts'].get('next_dist_obs'):
        assert width == 5 * n * m + n                       # connected | dist | dr | dr_total | next_dist | ues_at_bs
    else:
        assert width == 2 * n * m


@pytest.mark.parametrize('name', pending_movement_names())
def test_python_oracle_uniform_movement_matches_reference(name):
    """UniformMovement (util/movement.py:26-80) next to RandomWaypoint UEs: constant step, both components flip when the
    next point would not be strictly inside the map; two episodes ('slow' / 'fast' steps are redrawn at the reset).
    Oracle only -- the CUDA path does not offer this movement yet."""
    cfg, z = load_golden(name)
    env = po.OracleEnv(**oracle_kwargs(cfg))
    check_against_golden(env, cfg, z, exact_floats=True)
    mv = z['step_movement']                                     # [T, N, 5]: uniform UEs are (move_x, move_y, -1, 0, 0)
    uni = [i for i, u in enumerate(cfg['uniform_moves']) if u is not None]
    assert (mv[:, uni, 2] == -1).all()
    # at least one bounce happened: the sign of some UE's step changed within an episode
    assert any((np.sign(mv[:cfg['steps'], i, 0]) != np.sign(mv[0, i, 0])).any() or
               (np.sign(mv[:cfg['steps'], i, 1]) != np.sign(mv[0, i, 1])).any() for i in uni)
    W, H = cfg['map_wh']
    pos = z['step_pos'][:, uni]
    assert (pos[..., 0] > -21).all() and (pos[..., 0] < W + 21).all()


@pytest.mark.parametrize('name', pending_sequential_names())
def test_python_oracle_sequential_multi_agent_matches_reference(name):
    """SeqMultiAgentMobileEnv (multi_ue/multi_agent.py:110-179): one UE acts per call, the UEs move and time advances
    after the last one; observation row and multi-agent reward of the next UE.  Oracle only -- no CUDA path yet."""
    cfg, z = load_golden(name)
    env = po.OracleEnv(**oracle_kwargs(cfg))
    check_against_golden(env, cfg, z, exact_floats=True)
    n, m = cfg['n_ue'], len(cfg['bs_xy'])
    assert z['step_obs'].shape == (cfg['steps'], 4 * m + 1) and z['step_reward'].shape == (cfg['steps'],)
    assert z['step_time'][-1] == cfg['steps'] // n              # time advances once per round of the UEs
    moved = np.flatnonzero(np.diff(np.concatenate([[0], z['step_time']])))
    assert (moved % n == n - 1).all()


@pytest.mark.parametrize('kind', ['central', 'multi'])
def test_c_oracle_interference_extension_against_a_numpy_restatement(kind):
    """The interference extension has no counterpart in the reference (SNR only, station.py:122-127); its oracle is the C
    restatement.  Pin that restatement from a second side: numpy straight from the formula of docs/model.md with the
    interference term added -- SINR_b = P_b / (noise + sum of the other P_b') -- on the positions of the C trace, incl. a UE on
    top of a BS; and everything that follows from it in one step (in-range set, masks after the drop, unshared rates)."""
    from oracle.deepcomp_oracle import grid_layout
    n_ue, n_bs = 14, 9
    W, H, bs = grid_layout(n_bs)
    init_pos = [(bs[0][0], bs[0][1]), (bs[4][0] + 0.5, bs[4][1])] + [('random', 'random')] * (n_ue - 2)
    velocities = [0, 0] + ['slow'] * (n_ue - 2)
    kw = dict(kind=kind, n_ue=n_ue, bs_xy=bs, map_wh=(W, H), sharing='resource-fair', velocities=velocities, reward='avg',
              episode_length=30, init_pos=init_pos, seed=5)
    env = c_oracle.COracleEnv(interference=True, **kw)
    plain = c_oracle.COracleEnv(**kw)
    env.reset_trace(); plain.reset_trace()
    bsa = np.asarray(bs, dtype=np.float64)
    # station.py:26-30, 110-127 with numpy's own log10 / power
    ch = 0.8 + (1.1 * np.log10(2500.0) - 0.7) * 1.5 - 1.56 * np.log10(2500.0)
    c1 = 69.55 + 26.16 * np.log10(2500.0) - 13.82 * np.log10(50.0) - ch
    c2 = 44.9 - 6.55 * np.log10(50.0)
    rng = np.random.default_rng(2)
    saw_difference = 0.0
    for t in range(25):
        a = rng.integers(0, n_bs + 1, n_ue).astype(np.int32)
        w, p = env.step(a), plain.step(a)
        d = np.sqrt(((w['pos'][:, None, :] - bsa[None, :, :]) ** 2).sum(-1))
        power = 10.0 ** ((30.0 - (c1 + c2 * np.log10(d + 1e-16))) / 10.0)               # received power [mW]
        noise = 1e-9
        # (the others summed explicitly: `total - own` cancels catastrophically for a UE on top of a BS)
        others = np.stack([np.delete(power, b, axis=1).sum(1) for b in range(n_bs)], axis=1)
        sinr = power / (noise + others)
        np.testing.assert_allclose(w['snr'], sinr, rtol=1e-9, atol=0)
        # links exist only where the SINR clears the threshold, and a linked pair's unshared rate is bw log2(1 + SINR);
        # with resource-fair sharing the shared rate is that over the number of UEs at the BS (station.py:129-138, 173)
        assert not np.any(w['mask'].astype(bool) & ~(sinr > 2e-8))
        cnt = w['mask'].sum(0)
        exp_rate = np.where(w['mask'].astype(bool), 9e6 * np.log2(1.0 + sinr) / np.maximum(cnt, 1)[None, :], 0.0)
        np.testing.assert_allclose(w['link_rates'], exp_rate, rtol=1e-9, atol=0)
        assert np.array_equal(w['pos'], p['pos'])                                        # movement does not see the radio model
        saw_difference = max(saw_difference, float(np.max(np.abs(w['snr'] / p['snr'] - 1.0))))
    assert saw_difference > 1e-9
