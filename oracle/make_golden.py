"""TEST INFRASTRUCTURE ONLY -- freeze traces of the *unmodified reference env* into tests/golden/*.npz.

Run in a container that has /root/reference:   python -m oracle.make_golden
The reference (deepcomp.env.multi_ue.{central,multi_agent}) is imported under oracle/ref_stubs.py and driven with
seeded uniform-random actions; every array the parity tests compare is recorded per step.  The fixtures are the
pin that travels to the GPU box (where /root/reference does not exist).
"""
import json
import os
import sys

import numpy as np

from . import ref_loader as rl

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

STEP_KEYS = ['pos', 'mask', 'link_rates', 'snr', 'curr_dr', 'ewma', 'utility', 'movement', 'obs', 'reward',
             'lost_conn', 'sum_utility', 'time', 'num_ue']
RESET_KEYS = ['pos', 'mask', 'link_rates', 'snr', 'curr_dr', 'ewma', 'utility', 'movement', 'obs', 'num_ue']


def medium_map(bs_dist=100, border=10):
    """util/env_setup.py:87-104 create_dyn_medium_map -- BASELINE.json configs[0] (3 BS)"""
    y_dist = np.sqrt(bs_dist ** 2 - (bs_dist / 2) ** 2)
    wh = (2 * border + bs_dist, 2 * border + y_dist)
    bs = [(border, border), (border + bs_dist, border), (border + bs_dist / 2, border + y_dist)]
    return wh, bs


def scenarios():
    wh, bs = medium_map()
    S = []
    for kind in ('central', 'multi'):
        S.append(dict(name=f'medium3bs_5ue_{kind}', kind=kind, n_ue=5, bs_xy=bs, map_wh=wh, sharing='mixed',
                      velocities='slow', seed=42, reward='avg', steps=100, action_seed=1, episodes=2))
    W, H, gbs = rl.grid_layout(10)
    for kind in ('central', 'multi'):
        S.append(dict(name=f'grid10bs_50ue_{kind}', kind=kind, n_ue=50, bs_xy=gbs, map_wh=(W, H), sharing='mixed',
                      velocities='slow', seed=1000, reward='avg', steps=25, action_seed=2, episodes=1))
    W, H, gbs = rl.grid_layout(5)
    vel = ['slow', 'fast', 0, 2.5] * 3
    for sharing in ('resource-fair', 'rate-fair', 'proportional-fair', 'max-cap', 'mixed'):
        for kind, reward in (('central', 'avg'), ('multi', 'avg'), ('multi', 'min'), ('central', 'min'),
                             ('central', 'sum'), ('multi', 'sum')):
            if reward != 'avg' and sharing not in ('mixed', 'max-cap'):
                continue
            S.append(dict(name=f'grid5bs_12ue_{kind}_{reward}_{sharing}', kind=kind, n_ue=12, bs_xy=gbs,
                          map_wh=(W, H), sharing=sharing, velocities=vel, seed=7, reward=reward, steps=60,
                          action_seed=3, episodes=1))
    # fixed initial positions + long run on a tiny map: many waypoint redraws / pauses (movement.py:168-181)
    S.append(dict(name='tiny_redraw_multi', kind='multi', n_ue=6, bs_xy=[(30, 30), (90, 90)], map_wh=(120, 120),
                  sharing='mixed', velocities=['fast', 'fast', 'slow', 7, 'fast', 0], seed=99, reward='avg', steps=150,
                  action_seed=4, episodes=1, init_pos=[(60, 60), ('random', 'random'), (10, 110), (0, 0),
                                                        ('random', 5), (120, 120)]))
    return S


def population_scenarios():
    """Variable UE population: `ue_arrival` sequences of util/env_setup.py:205-226 and `new_ue_interval`
    (single_ue/base.py:433-443, 592-617) on envs with `max_ues` > `num_ue` (zero padding, multi_ue/central.py:46-55)."""
    S = []
    W, H, gbs = rl.grid_layout(5)
    base = dict(n_ue=3, bs_xy=gbs, map_wh=(W, H), sharing='mixed', velocities='slow', seed=33, steps=100,
                action_seed=9, episodes=2)
    largeupdown = {20: 1, 30: -1, 40: 1, 45: 1, 50: 1, 55: 2, 60: 3, 65: 2, 70: 1, 75: -1, 80: -2, 85: -3, 90: -3, 95: -2}
    for kind, reward in (('central', 'avg'), ('multi', 'avg'), ('multi', 'sum'), ('central', 'min')):
        S.append(dict(base, name=f'pop_largeupdown_{kind}_{reward}', kind=kind, reward=reward, max_ues=14,
                      ue_arrival=largeupdown))
    S.append(dict(base, name='pop_updown_multi_min', kind='multi', reward='min', max_ues=8,
                  ue_arrival={10: 1, 15: 1, 20: 1, 40: 1, 50: -1, 60: -1}))
    S.append(dict(base, name='pop_3up2down_central_sum', kind='central', reward='sum', max_ues=6, n_ue=2,
                  ue_arrival={10: 3, 30: -2}, sharing='proportional-fair'))
    S.append(dict(base, name='pop_interval25_multi_avg', kind='multi', reward='avg', max_ues=6, new_ue_interval=25))
    # UEs of different velocity classes change list positions when others leave (the arrivals are always 'slow')
    S.append(dict(base, name='pop_mixedvel_multi_avg', kind='multi', reward='avg', n_ue=4, max_ues=6,
                  velocities=['fast', 0, 'slow', 2.5], ue_arrival={8: 2, 20: -3, 30: 2, 45: -2}))
    S.append(dict(base, name='pop_mixedvel_central_avg', kind='central', reward='avg', n_ue=4, max_ues=6,
                  velocities=['fast', 0, 'slow', 2.5], ue_arrival={8: 2, 20: -3, 30: 2, 45: -2}))
    return S


def utility_scenarios():
    """User.util_func = 'step' (CLI --util step; user.py:81-92, env/util/utility.py:23-33)"""
    S = []
    W, H, gbs = rl.grid_layout(5)
    for kind, reward in (('central', 'avg'), ('multi', 'avg'), ('multi', 'min'), ('multi', 'sum')):
        S.append(dict(name=f'utilstep_{kind}_{reward}', kind=kind, n_ue=12, bs_xy=gbs, map_wh=(W, H), sharing='mixed',
                      velocities='slow', seed=17, reward=reward, steps=50, action_seed=6, episodes=1, util_func='step'))
    return S


def obs_variant_scenarios():
    """MaxNormEnv observation (single_ue/variants.py:308-332): CentralMaxNormEnv (multi_ue/central.py:155-164) and the
    same composition over MultiAgentMobileEnv (oracle/ref_loader.py:build_env).  UE 12 starts on top of BS 0 (d = 0:
    SNR far above the cap), UE 11 in the far corner (every BS out of range: all entries negative)."""
    S = []
    W, H, gbs = rl.grid_layout(5)
    vel = ['slow', 'fast', 0, 2.5] * 3
    init = [('random', 'random')] * 10 + [(W, H), gbs[0]]
    for kind, reward in (('central', 'avg'), ('multi', 'avg')):
        S.append(dict(name=f'maxnorm_{kind}_{reward}', kind=kind, n_ue=12, bs_xy=gbs, map_wh=(W, H), sharing='mixed',
                      velocities=vel, seed=23, reward=reward, steps=50, action_seed=8, episodes=1, obs_norm='max',
                      init_pos=init))
    S.append(dict(name='pop_maxnorm_central_sum', kind='central', reward='sum', n_ue=2, max_ues=6, bs_xy=gbs,
                  map_wh=(W, H), sharing='mixed', velocities='slow', seed=33, steps=60, action_seed=9, episodes=1,
                  ue_arrival={10: 3, 30: -2}, obs_norm='max'))
    return S


def pending_obs_scenarios():
    """Observation classes the oracle restates but the CUDA path does not offer yet: CentralNormDrEnv
    (central.py:107-140 over variants.py:173-250) and CentralDrEnv (central.py:75-104 over variants.py:42-170) with its
    env_config options.  Traces for the next round's kernels; tests/test_oracle_golden.py pins the oracle on them."""
    S = []
    W, H, gbs = rl.grid_layout(5)
    vel = ['slow', 'fast', 0, 2.5] * 3
    base = dict(kind='central', n_ue=12, bs_xy=gbs, map_wh=(W, H), velocities=vel, seed=29, reward='avg', steps=50,
                action_seed=12, episodes=1)
    S.append(dict(base, name='normdr_central_mixed', sharing='mixed', obs_variant='normdr'))
    S.append(dict(base, name='normdr_central_max-cap', sharing='max-cap', obs_variant='normdr'))
    S.append(dict(base, name='datarate_central_auto_all', sharing='mixed', obs_variant='datarate', util_func='step',
                  obs_opts=dict(dr_cutoff='auto', sub_req_dr=True, curr_dr_obs=True, ues_at_bs_obs=True, dist_obs=True,
                                next_dist_obs=True)))
    S.append(dict(base, name='datarate_central_cutoff200', sharing='mixed', obs_variant='datarate',
                  obs_opts=dict(dr_cutoff=200, sub_req_dr=False)))
    # UniformMovement (util/movement.py:26-80): constant step, both components flip when the next point would not be
    # strictly inside the map; mixed with RandomWaypoint UEs; long enough for several bounces
    um = [('slow', 'slow'), ('fast', 'slow'), None, (3, -2), (0, 7.5), None, ('fast', 'fast'), (-4, 0)]
    for kind in ('central', 'multi'):
        S.append(dict(name=f'uniform_{kind}_avg', kind=kind, n_ue=8, bs_xy=gbs, map_wh=(W, H), sharing='mixed',
                      velocities='slow', seed=31, reward='avg', steps=120, action_seed=13, episodes=2,
                      uniform_moves=um))
    # SeqMultiAgentMobileEnv (multi_ue/multi_agent.py:110-179): one UE acts per call; 6 UEs x 40 time steps
    for reward in ('avg', 'min'):
        S.append(dict(name=f'seq_multi_{reward}', kind='multi', n_ue=6, bs_xy=gbs, map_wh=(W, H), sharing='mixed',
                      velocities=['slow', 'fast', 0, 2.5, 'slow', 'fast'], seed=37, reward=reward, steps=240,
                      action_seed=14, episodes=1, sequential=True))
    return S


def brute_scenarios():
    """BruteForceAgent (deepcomp/agent/brute_force.py:59-94): all (M+1)^N joint actions tested with
    MobileEnv.test_ue_actions (single_ue/base.py:284-313) on the central env, the best one taken."""
    S = []
    wh, bs = medium_map()
    for sharing, reward, n_ue in (('mixed', 'avg', 4), ('proportional-fair', 'sum', 3), ('max-cap', 'min', 3),
                                  ('rate-fair', 'avg', 5)):
        S.append(dict(name=f'brute_{sharing}_{reward}_{n_ue}ue', kind='central', n_ue=n_ue, bs_xy=bs, map_wh=wh,
                      sharing=sharing, velocities='slow', seed=5, reward=reward, steps=12, action_seed=0, episodes=1))
    return S


def record_brute(sc):
    """Per step: the reward of EVERY candidate action (brute_force.py:64-77) from the state before the step, then the
    step with the best one (np.argmax: first maximum, brute_force.py:90-92)."""
    env = rl.build_env(sc['kind'], sc['n_ue'], sc['seed'], sc['bs_xy'], sc['map_wh'], sharing=sc['sharing'],
                       velocities=sc['velocities'], reward=sc['reward'], episode_length=sc['steps'])
    tr = rl.RefTrace(env, sc['kind'])
    n, m = sc['n_ue'], len(sc['bs_xy'])
    n_cand = (m + 1) ** n
    cand_rewards, actions, steps = [], [], {k: [] for k in STEP_KEYS}
    env.reset()
    reset = tr.snapshot()
    for t in range(sc['steps']):
        rew = np.zeros(n_cand)
        for c in range(n_cand):
            digits = np.base_repr(c, base=m + 1).zfill(n)                      # == number_to_base(c, m + 1, n)
            act = [int(ch, m + 1) for ch in digits]
            rewards = env.test_ue_actions(env.get_ue_actions(act))
            rew[c] = float(env.step_reward(rewards))
        best = int(np.argmax(rew))
        a = np.array([int(ch, m + 1) for ch in np.base_repr(best, base=m + 1).zfill(n)], dtype=np.int32)
        cand_rewards.append(rew)
        actions.append(a)
        s = tr.step(a)
        for k in STEP_KEYS:
            steps[k].append(s[k])
    out = {'cand_rewards': np.stack(cand_rewards), 'actions': np.stack(actions)}
    for k in STEP_KEYS:
        out['step_' + k] = np.stack([np.asarray(v) for v in steps[k]])
    for k in ('pos', 'mask', 'ewma'):
        out['reset_' + k] = reset[k]
    cfg = dict(sc)
    cfg['bs_xy'] = [[float(x), float(y)] for x, y in sc['bs_xy']]
    cfg['map_wh'] = [float(sc['map_wh'][0]), float(sc['map_wh'][1])]
    out['config'] = np.array(json.dumps(cfg))
    return out


def policy_scenarios():
    """Closed loops of the reference's baseline agents (deepcomp/agent/heuristics.py, dummy.py) on the reference env."""
    S = []
    W, H, gbs = rl.grid_layout(5)
    vel = ['slow', 'fast', 0, 2.5] * 3
    base = dict(kind='multi', n_ue=12, bs_xy=gbs, map_wh=(W, H), sharing='mixed', velocities=vel, seed=21, reward='avg',
                steps=60, action_seed=0, episodes=2)
    S.append(dict(base, name='policy_3gpp', policy=dict(kind='3gpp')))
    S.append(dict(base, name='policy_fullcomp', policy=dict(kind='fullcomp')))
    S.append(dict(base, name='policy_dynamic_eps05', policy=dict(kind='dynamic', epsilon=0.5)))
    S.append(dict(base, name='policy_dynamic_eps001', policy=dict(kind='dynamic', epsilon=0.01)))
    S.append(dict(base, name='policy_static_c2', policy=dict(kind='static', cluster_size=2, seed=5)))
    W, H, gbs = rl.grid_layout(10)
    big = dict(kind='multi', n_ue=50, bs_xy=gbs, map_wh=(W, H), sharing='mixed', velocities='slow', seed=1000,
               reward='avg', steps=25, action_seed=0, episodes=1)
    S.append(dict(big, name='policy_3gpp_grid10', policy=dict(kind='3gpp')))
    # StaticClustering ranks the cells of a cluster with a stable sort over a Python set of Basestation objects
    # (heuristics.py:171-183): equal-dr ties resolve in hash (= memory address) order, i.e. nondeterministically.  Integer
    # start positions on a regular grid produce such ties, so this trace starts the UEs at non-integer positions.
    frac = np.random.default_rng(3)
    init = [(round(float(frac.uniform(0, W)), 3), round(float(frac.uniform(0, H)), 3)) for _ in range(50)]
    S.append(dict(big, name='policy_static_c3_grid10', policy=dict(kind='static', cluster_size=3, seed=1),
                  init_pos=init))
    S.append(dict(base, name='policy_fixed', kind='central', policy=dict(kind='fixed', action=[1, 2, 3, 4, 5, 0] * 2,
                                                                        noop_interval=2)))
    return S


def make_reference_agent(sc, env):
    from deepcomp.agent.heuristics import Heuristic3GPP, FullCoMP, DynamicSelection, StaticClustering
    from deepcomp.agent.dummy import FixedAgent
    pol = sc['policy']
    if pol['kind'] == '3gpp':
        return Heuristic3GPP()
    if pol['kind'] == 'fullcomp':
        return FullCoMP()
    if pol['kind'] == 'dynamic':
        return DynamicSelection(epsilon=pol['epsilon'])
    if pol['kind'] == 'static':
        return StaticClustering(cluster_size=pol['cluster_size'], bs_list=env.bs_list, seed=pol['seed'])
    if pol['kind'] == 'fixed':
        return FixedAgent(action=np.array(pol['action']), noop_interval=pol['noop_interval'])
    raise ValueError(pol)


def record(sc):
    env = rl.build_env(sc['kind'], sc['n_ue'], sc['seed'], sc['bs_xy'], sc['map_wh'], sharing=sc['sharing'],
                       velocities=sc['velocities'], reward=sc['reward'], episode_length=sc['steps'],
                       init_pos=sc.get('init_pos'), max_ues=sc.get('max_ues'), ue_arrival=sc.get('ue_arrival'),
                       new_ue_interval=sc.get('new_ue_interval'), util_func=sc.get('util_func', 'log'),
                       obs_norm=sc.get('obs_norm', 'rel'), obs_variant=sc.get('obs_variant'),
                       obs_opts=sc.get('obs_opts'), uniform_moves=sc.get('uniform_moves'),
                       sequential=sc.get('sequential', False))
    pop = 'max_ues' in sc
    n_act = sc.get('max_ues', sc['n_ue'])                # length of the action vector (central.py:28)
    tr = rl.RefTrace(env, sc['kind'])
    rng = np.random.default_rng(sc['action_seed'])
    n_bs = len(sc['bs_xy'])
    out = {}
    actions = []
    steps = {k: [] for k in STEP_KEYS}
    resets = {k: [] for k in RESET_KEYS}
    agent = make_reference_agent(sc, env) if 'policy' in sc else None
    for ep in range(sc['episodes']):
        raw_obs = env.reset()
        r = tr.snapshot()
        r['obs'] = tr.flat_obs(raw_obs)
        for k in RESET_KEYS:
            resets[k].append(r[k])
        for t in range(sc['steps']):
            if agent is None:
                a = rng.integers(0, n_bs + 1, n_act).astype(np.int32)
            elif getattr(agent, 'central_agent', False):      # StaticClustering never calls super().__init__()
                a = np.asarray(agent.compute_action(raw_obs), dtype=np.int32)           # simulation.py:327-349
            else:
                # simulation.py:351-380: one compute_action call per agent id
                a = np.array([agent.compute_action(raw_obs[ue.id], policy_id='ue') for ue in env.ue_list],
                             dtype=np.int32)
            s = tr.step(a)
            raw_obs = tr.last_obs
            # base.py:371-381: done is None; multi_agent.py:97-102: dict of None incl. '__all__'
            if sc['kind'] == 'central':
                assert s['done'] is None
            elif sc.get('sequential'):
                # multi_agent.py:134-141 wraps MultiAgentMobileEnv.done() (already a dict) under the two keys
                assert '__all__' in s['done'] and len(s['done']) == 2
                assert all(v is None for d in s['done'].values() for v in d.values())
            elif pop:
                assert all(v is None for v in s['done'].values())
            else:
                assert set(s['done'].keys()) == {str(i + 1) for i in range(sc['n_ue'])} | {'__all__'}
                assert all(v is None for v in s['done'].values())
            actions.append(a)
            for k in STEP_KEYS:
                steps[k].append(s[k])
    out['actions'] = np.stack(actions)
    for k in STEP_KEYS:
        out['step_' + k] = np.stack([np.asarray(v) for v in steps[k]])
    for k in RESET_KEYS:
        out['reset_' + k] = np.stack([np.asarray(v) for v in resets[k]])
    if agent is not None and sc['policy']['kind'] == 'static':
        masks = np.zeros(n_bs, dtype=np.uint64)
        for bs, members in agent.clusters.items():
            for m in members:
                masks[env.bs_list.index(bs)] |= np.uint64(1) << np.uint64(env.bs_list.index(m))
        out['cluster_masks'] = masks
    cfg = {k: v for k, v in sc.items()}
    if cfg.get('ue_arrival') is not None:
        cfg['ue_arrival'] = {str(t): int(n) for t, n in cfg['ue_arrival'].items()}
    cfg['bs_xy'] = [[float(x), float(y)] for x, y in sc['bs_xy']]
    cfg['map_wh'] = [float(sc['map_wh'][0]), float(sc['map_wh'][1])]
    out['config'] = np.array(json.dumps(cfg))
    return out


def anchors():
    """Known answers that exist only as comments in the reference, evaluated by the
    reference's own functions."""
    R = rl._import_reference()
    from deepcomp.env.util.utility import log_utility
    import random
    bs = R['Basestation']('A', R['Point'](0, 0), 'resource-fair')
    d = np.array([0.0, 1e-3, 0.5, 1, 10, 11, 12.16, 46, 68, 68.9, 68.92488308058006, 68.92488308058007, 69, 100, 250, 1000])
    snr = np.array([float(bs.snr(R['Point'](x, 0))) for x in d])
    ue = type('U', (), {})()
    rate = []
    for x in d:
        ue.pos = R['Point'](x, 0)
        rate.append(float(bs.data_rate_unshared(ue)))
    dr = np.array([0, 1e-9, 0.005, 0.01, 0.5, 1, 2.5, 50, 99.9, 100, 1e4])
    util = np.array([float(log_utility(x)) for x in dr])
    seeds = [142, 1100, 0, 2 ** 32 + 5, 7 + 100 * 12]
    raw = np.array([[random.Random(s).getrandbits(32)] + [0] * 7 for s in seeds], dtype=np.uint64)
    for i, s in enumerate(seeds):
        r = random.Random(s)
        raw[i] = [r.getrandbits(32) for _ in range(8)]
    ranges = [(0, 300), (0, 300), (1, 3), (10, 290), (10, 290), (5, 10), (0, 120), (10, 110)] * 8
    ints = np.zeros((len(seeds), len(ranges)), dtype=np.int64)
    for i, s in enumerate(seeds):
        r = random.Random(s)
        ints[i] = [r.randint(a, b) for a, b in ranges]
    return dict(dist=d, snr=snr, rate_unshared=np.array(rate), dr=dr, log_utility=util,
                path_loss_1m=np.float64(bs.path_loss(1)), rng_seeds=np.array(seeds, dtype=np.int64), rng_raw=raw,
                rng_ranges=np.array(ranges, dtype=np.int64), rng_ints=ints)


def main():
    if not rl.available():
        sys.exit('reference not found under ' + rl.REFERENCE_ROOT)
    os.makedirs(OUT_DIR, exist_ok=True)
    only = sys.argv[1] if len(sys.argv) > 1 else ''     # optional name prefix: regenerate a subset only
    if not only:
        np.savez_compressed(os.path.join(OUT_DIR, 'anchors.npz'), **anchors())
    for sc in scenarios() + policy_scenarios() + population_scenarios() + brute_scenarios() + utility_scenarios() + \
            obs_variant_scenarios() + pending_obs_scenarios():
        if not sc['name'].startswith(only):
            continue
        data = record_brute(sc) if sc['name'].startswith('brute_') else record(sc)
        path = os.path.join(OUT_DIR, sc['name'] + '.npz')
        np.savez_compressed(path, **data)
        print(f"{sc['name']}: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == '__main__':
    main()
