"""CPU: the host form of the baseline agents against action traces of the reference's own agents (tests/golden/policy_*)."""
import numpy as np
import pytest

from deepcomp_b200 import agents, sharing_for_bs
from deepcomp_b200.entities import Basestation, Point

from helpers import golden_names, load_golden

POLICY_CASES = [n for n in golden_names() if n.startswith('policy_')]


def make_agent(cfg, z=None):
    pol = cfg['policy']
    if pol['kind'] == '3gpp':
        return agents.Heuristic3GPP()
    if pol['kind'] == 'fullcomp':
        return agents.FullCoMP()
    if pol['kind'] == 'dynamic':
        return agents.DynamicSelection(epsilon=pol['epsilon'])
    if pol['kind'] == 'static':
        bs = [Basestation(str(b), Point(x, y), sharing_for_bs(cfg['sharing'], b)) for b, (x, y) in enumerate(cfg['bs_xy'])]
        return agents.StaticClustering(cluster_size=pol['cluster_size'], bs_list=bs, seed=pol['seed'])
    if pol['kind'] == 'fixed':
        return agents.FixedAgent(action=np.array(pol['action']), noop_interval=pol['noop_interval'])
    raise ValueError(pol)


def obs_dicts(flat, n_bs):
    """golden multi-agent obs [N, 4M+1] (alphabetical key order) -> list of per-UE obs dicts"""
    m = n_bs
    return [{'connected': [int(v) for v in row[:m]], 'dr': list(row[m:2 * m]), 'ues_at_bs': list(row[2 * m:3 * m]),
             'util_at_bs': list(row[3 * m:4 * m]), 'utility': [row[4 * m]]} for row in flat]


@pytest.mark.parametrize('name', POLICY_CASES)
def test_host_agents_reproduce_reference_actions(name):
    cfg, z = load_golden(name)
    agent = make_agent(cfg)
    n_bs = len(cfg['bs_xy'])
    t = 0
    for ep in range(cfg['episodes']):
        obs = z['reset_obs'][ep]
        for _ in range(cfg['steps']):
            if cfg['policy']['kind'] == 'fixed':
                a = np.asarray(agent.compute_action(None), dtype=np.int32)
            else:
                a = np.array([agent.compute_action(o, 'ue') for o in obs_dicts(obs, n_bs)], dtype=np.int32)
            assert np.array_equal(a, z['actions'][t]), (name, t)
            obs = z['step_obs'][t]
            t += 1


@pytest.mark.parametrize('name', [n for n in POLICY_CASES if 'fixed' not in n])
def test_batched_agents_reproduce_reference_actions(name):
    """compute_actions over the packed observation of all UEs (and of a stack of steps) at once == the reference's agents"""
    cfg, z = load_golden(name)
    agent = make_agent(cfg)
    obs = np.concatenate([z['reset_obs'][:1], z['step_obs'][:cfg['steps'] - 1]])       # obs seen before each action
    got = agent.compute_actions(obs).numpy()
    assert np.array_equal(got, z['actions'][:cfg['steps']])
    assert np.array_equal(agent.compute_actions(obs[3]).numpy(), z['actions'][3])


@pytest.mark.parametrize('name', [n for n in POLICY_CASES if 'static' in n])
def test_static_clusters_match_reference(name):
    """build_clusters (heuristics.py:132-167) with the same seeded random.Random picks the same clusters"""
    cfg, z = load_golden(name)
    assert np.array_equal(make_agent(cfg).cluster_masks(), z['cluster_masks'])


def test_device_policy_specs():
    assert agents.Heuristic3GPP().device_policy() == {'kind': '3gpp'}
    assert agents.DynamicSelection(0.25).device_policy() == {'kind': 'dynamic', 'epsilon': 0.25}
    with pytest.raises(AssertionError):
        agents.DynamicSelection(1.5)                                   # cli.py:100
    spec = agents.FixedAgent([1, 0, 2], noop_interval=3).device_policy()
    assert spec['kind'] == 'fixed' and spec['noop_interval'] == 3 and spec['fixed_action'].tolist() == [1, 0, 2]
    assert spec['calls_before'] == 0
