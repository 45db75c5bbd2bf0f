"""CPU, this container only: lock-step of the Python restatement against the live, unmodified reference env."""
import numpy as np
import pytest

from oracle import ref_loader as rl
from oracle import deepcomp_oracle as po

from helpers import assert_exact

pytestmark = pytest.mark.skipif(not rl.available(), reason='/root/reference not present (GPU box)')

KEYS = ['pos', 'mask', 'link_rates', 'snr', 'curr_dr', 'ewma', 'utility', 'movement', 'obs']


@pytest.mark.parametrize('kind,reward', [('central', 'avg'), ('multi', 'avg'), ('multi', 'min'), ('central', 'sum')])
@pytest.mark.parametrize('seed', [3, 20240])
def test_lockstep_random_scenarios(kind, reward, seed):
    rng = np.random.default_rng(seed)
    n_bs = int(rng.integers(1, 8))
    n_ue = int(rng.integers(1, 15))
    W, H, bs = po.grid_layout(n_bs)
    sharing = [['resource-fair', 'rate-fair', 'proportional-fair', 'max-cap'][int(rng.integers(0, 4))]
               for _ in range(n_bs)]
    vel = [['slow', 'fast', 0, 4][int(rng.integers(0, 4))] for _ in range(n_ue)]
    ref = rl.RefTrace(rl.build_env(kind, n_ue, seed, bs, (W, H), sharing=sharing, velocities=vel, reward=reward,
                                   episode_length=40), kind)
    orc = po.OracleEnv(kind, n_ue, bs, (W, H), sharing=sharing, velocities=vel, seed=seed, reward=reward,
                       episode_length=40)
    for ep in range(2):
        a, b = ref.reset(), orc.reset_trace()
        for k in KEYS:
            assert_exact(b[k], a[k], f'reset.{k}')
        for t in range(40):
            act = rng.integers(0, n_bs + 1, n_ue)
            a, b = ref.step(act), orc.step(act)
            for k in KEYS + ['reward', 'lost_conn', 'sum_utility', 'time']:
                assert_exact(b[k], a[k], f'step[{t}].{k}')


def test_rand_episodes_continue_the_rng_stream():
    """rand_episodes=True: no reseed on reset (base.py:171-173) -> second episode differs, both match."""
    W, H, bs = po.grid_layout(4)
    ref = rl.RefTrace(rl.build_env('central', 6, 11, bs, (W, H), rand_episodes=True, episode_length=10), 'central')
    orc = po.OracleEnv('central', 6, bs, (W, H), seed=11, rand_episodes=True, episode_length=10)
    first = None
    for ep in range(3):
        a, b = ref.reset(), orc.reset_trace()
        assert_exact(b['pos'], a['pos'], 'pos')
        assert_exact(b['movement'], a['movement'], 'movement')
        if first is None:
            first = a['pos']
        else:
            assert not np.array_equal(first, a['pos'])
        for t in range(10):
            act = np.zeros(6, dtype=int)
            a, b = ref.step(act), orc.step(act)
            assert_exact(b['pos'], a['pos'], 'pos')


@pytest.mark.parametrize('variant', ['maxnorm', 'normdr', 'datarate', 'uniform', 'sequential'])
@pytest.mark.parametrize('seed', [5, 777])
def test_lockstep_observation_movement_and_stepping_variants(variant, seed):
    """The classes beyond RelNorm / RandomWaypoint (MaxNormEnv, CentralNormDrEnv, CentralDrEnv with random options,
    UniformMovement, SeqMultiAgentMobileEnv) on random scenarios: oracle == live reference, bit for bit."""
    rng = np.random.default_rng(seed)
    n_bs = int(rng.integers(2, 7))
    n_ue = int(rng.integers(2, 10))
    W, H, bs = po.grid_layout(n_bs)
    sharing = [['resource-fair', 'rate-fair', 'proportional-fair', 'max-cap'][int(rng.integers(0, 4))]
               for _ in range(n_bs)]
    vel = [['slow', 'fast', 0, 4][int(rng.integers(0, 4))] for _ in range(n_ue)]
    kind, extra = 'central', {}
    if variant == 'maxnorm':
        kind = ['central', 'multi'][int(rng.integers(0, 2))]
        extra = dict(obs_norm='max')
    elif variant == 'normdr':
        extra = dict(obs_variant='normdr')
    elif variant == 'datarate':
        auto = bool(rng.integers(0, 2))
        sub = True if auto else bool(rng.integers(0, 2))
        dist_obs = bool(rng.integers(0, 2))
        extra = dict(obs_variant='datarate', util_func='step' if auto else 'log',
                     obs_opts=dict(dr_cutoff='auto' if auto else 250, sub_req_dr=sub,
                                   curr_dr_obs=auto and bool(rng.integers(0, 2)), ues_at_bs_obs=bool(rng.integers(0, 2)),
                                   dist_obs=dist_obs, next_dist_obs=dist_obs and bool(rng.integers(0, 2))))
    elif variant == 'uniform':
        kind = ['central', 'multi'][int(rng.integers(0, 2))]
        choices = [None, ('slow', 'slow'), ('fast', 'slow'), (3, -2), (0, 6.5), ('fast', 'fast')]
        extra = dict(uniform_moves=[choices[int(rng.integers(0, len(choices)))] for _ in range(n_ue)])
    else:
        kind = 'multi'
        extra = dict(sequential=True)
    reward = ['avg', 'min'][int(rng.integers(0, 2))]
    ref = rl.RefTrace(rl.build_env(kind, n_ue, seed, bs, (W, H), sharing=sharing, velocities=vel, reward=reward,
                                   episode_length=60, **extra), kind)
    orc = po.OracleEnv(kind, n_ue, bs, (W, H), sharing=sharing, velocities=vel, seed=seed, reward=reward,
                       episode_length=60, **extra)
    for ep in range(2):
        a, b = ref.reset(), orc.reset_trace()
        for k in KEYS:
            assert_exact(b[k], a[k], f'{variant}.reset.{k}')
        for t in range(60):
            act = rng.integers(0, n_bs + 1, n_ue)
            a, b = ref.step(act), orc.step(act)
            for k in KEYS + ['reward', 'lost_conn', 'sum_utility', 'time']:
                assert_exact(b[k], a[k], f'{variant}.step[{t}].{k}')
