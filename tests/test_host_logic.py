"""CPU: host-side logic of the product that needs no device (seeding, sharding, config parsing, spaces)."""
import numpy as np
import pytest

from deepcomp_b200 import env_seeds, sharing_for_bs, spaces
from deepcomp_b200.distributed import shard_bounds
from deepcomp_b200.entities import Basestation, Map, Point, RandomWaypoint, User, create_ues
from deepcomp_b200.env import _parse_env_config


def test_env_seeds_do_not_collide_across_envs():
    """UE i of env k draws from seed_k + 100*i (base.py:138-143): all K*N UE seeds must be distinct."""
    K, N = 200, 50
    s = env_seeds(1000, K, N)
    ue = (s[:, None] + 100 * np.arange(1, N + 1)[None, :]).ravel()
    assert len(np.unique(ue)) == K * N
    assert np.array_equal(env_seeds(1000, 8, N, first_env=4), env_seeds(1000, 12, N)[4:])


def test_sharing_mix_follows_reference():
    """env_setup.py:40-49"""
    assert [sharing_for_bs('mixed', b) for b in range(5)] == ['resource-fair', 'rate-fair', 'proportional-fair',
                                                              'resource-fair', 'rate-fair']
    assert sharing_for_bs('max-cap', 3) == 'max-cap'
    with pytest.raises(ValueError):
        sharing_for_bs('bogus', 0)


@pytest.mark.parametrize('total,world', [(8192, 8), (10, 3), (1, 4), (1024, 1)])
def test_shard_bounds_partition_the_batch(total, world):
    spans = [shard_bounds(total, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def _config(**over):
    m = Map(120, 106.6)
    bs = [Basestation('A', Point(10, 10), 'resource-fair'), Basestation('B', Point(110, 10), 'rate-fair')]
    ues = create_ues(m, 1, 2, 1)
    cfg = {'episode_length': 100, 'seed': 42, 'map': m, 'bs_list': bs, 'ue_list': ues, 'rand_episodes': False,
           'new_ue_interval': None, 'reward': 'avg', 'max_ues': None, 'ue_arrival': None, 'log_metrics': True,
           'dashboard': False, 'ue_details': False}
    cfg.update(over)
    return cfg


def test_env_config_parsing_matches_reference_fields():
    sc = _parse_env_config(_config())
    assert sc['n_ue'] == 4 and sc['map_wh'] == (120, 106)            # Map casts to int (map.py:20-21)
    assert sc['velocities'] == [0, 'slow', 'slow', 'fast']           # env_setup.py:145-161
    assert sc['sharing'] == ['resource-fair', 'rate-fair']
    assert sc['init_pos'] == [('random', 'random')] * 4
    assert sc['pause_duration'] == 2 and sc['border_buffer'] == 10   # movement.py:87
    assert [ue.id for ue in _config()['ue_list']] == ['1', '2', '3', '4']


def test_env_config_variable_population_fields():
    """max_ues as the reference derives it when None (base.py:80-84, 191-203): initial UEs + every arrival"""
    assert _parse_env_config(_config())['max_ues'] == 4
    assert _parse_env_config(_config(max_ues=7))['max_ues'] == 7
    sc = _parse_env_config(_config(new_ue_interval=10))
    assert sc['max_ues'] == 4 + 9 and sc['new_ue_interval'] == 10        # int((100 - 1) / 10)
    sc = _parse_env_config(_config(ue_arrival={10: 3, 30: -2}))
    assert sc['max_ues'] == 4 + 3 and sc['ue_arrival'] == {10: 3, 30: -2}
    assert _parse_env_config(_config(ue_arrival={10: 1, 20: -1, 30: 1, 40: -1}))['max_ues'] == 5
    with pytest.raises(AssertionError):
        _parse_env_config(_config(max_ues=3))                        # base.py:84


def test_env_config_out_of_scope_features_fail_loudly():
    cfg = _config()
    cfg['ue_list'][0].util_func = 'step'
    with pytest.raises(NotImplementedError):
        _parse_env_config(cfg)
    cfg = _config()
    del cfg['seed']
    with pytest.raises(KeyError):
        _parse_env_config(cfg)


def test_entities_validate_like_the_reference():
    with pytest.raises(AssertionError):
        Basestation('A', Point(0, 0), 'bogus')                       # station.py:21
    with pytest.raises(AssertionError):
        RandomWaypoint(Map(100, 100), 'slow', border_buffer=0)       # movement.py:103
    with pytest.raises(AssertionError):
        User('1', Map(100, 100), 0, 0, RandomWaypoint(Map(100, 100), 1), util_func='bogus')   # user.py:30


def test_spaces_contract():
    """What the facades / RLlib touch: contains, shape, alphabetical Dict order."""
    d = spaces.Dict({'utility': spaces.Box(low=-1, high=1, shape=(1,)), 'connected': spaces.MultiBinary(3),
                     'dr': spaces.Box(low=0, high=1, shape=(3,))})
    assert list(d.spaces.keys()) == ['connected', 'dr', 'utility']
    assert spaces.Discrete(4).contains(3) and not spaces.Discrete(4).contains(4)
    md = spaces.MultiDiscrete([4, 4, 4])
    assert md.contains(np.array([0, 3, 1])) and not md.contains(np.array([0, 4, 1]))
    assert md.shape == (3,)


def test_fixed_point_utility_sums_of_the_fx_agg_experiment():
    """Arithmetic of the -DDCB_FX_AGG experiment (dcb_step.cu), mirrored in numpy on the utilities of a
    reference trace: a utility as a 2^-36 fixed-point number, split into a non-negative 20-bit low part and a signed
    high part, each summed in int32 (native shared-memory atomics on the device) -- the sums stay inside int32 for 512
    UEs per BS, are independent of the order, and come back within |C_b| * 2^-37 of the exact sum (1e-9 parity bar)."""
    from helpers import load_golden
    cfg, z = load_golden('grid10bs_50ue_multi')
    util = z['step_utility'].astype(np.float64)                  # [T, N], values in [-20, 20]
    mask = z['step_mask'].astype(bool)                           # [T, N, M]
    fx = np.rint(util * 2.0 ** 36).astype(np.int64)
    lo, hi = fx & 0xfffff, fx >> 20
    assert (lo >= 0).all() and (lo < 2 ** 20).all() and (np.abs(hi) < 2 ** 21).all()
    assert np.array_equal((hi << 20) + lo, fx)
    worst = 0.0
    for t in range(util.shape[0]):
        for b in range(mask.shape[2]):
            m = mask[t, :, b]
            s_lo, s_hi = lo[t, m].astype(np.int32).sum(dtype=np.int32), hi[t, m].astype(np.int32).sum(dtype=np.int32)
            got = float((np.int64(s_hi) << 20) + np.int64(s_lo)) * 2.0 ** -36
            perm = np.random.default_rng(t * 16 + b).permutation(int(m.sum()))
            assert (lo[t, m][perm].sum(), hi[t, m][perm].sum()) == (int(s_lo), int(s_hi))
            want = float(np.sum(util[t, m]))
            assert abs(got - want) <= m.sum() * 2.0 ** -37 + 1e-13
            worst = max(worst, abs(got - want))
    assert worst < 1e-9
    # headroom: 512 UEs (the fused kernel's limit) at +-20 on one BS
    assert 512 * (2 ** 20 - 1) < 2 ** 31 and 512 * (20 * 2 ** 16) < 2 ** 31


def test_bench_arms_share_one_workload_config():
    """bench.py: the GPU arm and the reference arm print the same `config`; fragments never exceed the timed region, an
    episode, or the memory clamp; --total-envs splits the batch (strong scaling)"""
    import argparse
    import bench
    args = argparse.Namespace(n_ue=50, n_bs=10, envs=1024, kind='multi', sharing='mixed', episode_length=100, fragment=100,
                              steps=20, seed=1000)
    assert bench.effective_fragment(args) == 20
    cfg = bench.workload_config(args, 8)
    assert cfg == bench.workload_config(args, 8) and '8192 total' in cfg['workload'] and cfg['fragment_steps'] == 20
    assert 'l2' in cfg and '168 MB' in cfg['l2']
    args.steps, args.envs = 5000, 65536 * 16
    assert 1 <= bench.effective_fragment(args) < 100          # observation buffer of one fragment stays under the clamp
    args.envs, args.kind = 1024, 'central'
    assert bench.effective_fragment(args) == 100
    assert bench.obs_floats_per_step(args) == 1024 * (2 * 50 * 10 + 50)
