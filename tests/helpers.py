"""Shared helpers for the parity tests: golden-fixture loading, oracle construction, comparisons."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# bit-exact: integer state and positions; within tolerance: every float aggregate.
EXACT_KEYS = ['pos', 'mask', 'movement']
EXACT_STEP_KEYS = ['lost_conn', 'time']
FLOAT_KEYS = ['link_rates', 'snr', 'curr_dr', 'ewma', 'utility', 'obs']
FLOAT_STEP_KEYS = ['reward', 'sum_utility']

# north_star: SINR / data-rate / reward within 1e-5 relative.  The restatements are fp64 and land ~1e-13; the
# tests hold them to RTOL below (abs floor ATOL for quantities that cross zero, e.g. utility in dB).
RTOL = 1e-9
ATOL = 1e-9


def _names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz'))
                  if not p.endswith('anchors.npz'))


def golden_names():
    """fixed-population traces (every test that replays a trace step by step)"""
    return [n for n in _names() if not n.startswith(('pop_', 'brute_', 'utilstep_', 'maxnorm_', 'normdr_', 'datarate_',
                                                      'uniform_', 'seq_'))]


def pending_obs_names():
    """traces of the data-rate observation classes (CentralNormDrEnv, CentralDrEnv: central.py:75-140)"""
    return [n for n in _names() if n.startswith(('normdr_', 'datarate_'))]


def pending_sequential_names():
    """traces of SeqMultiAgentMobileEnv (multi_ue/multi_agent.py:110-179)"""
    return [n for n in _names() if n.startswith('seq_')]


def pending_movement_names():
    """traces with UniformMovement UEs (util/movement.py:26-80), mixed with RandomWaypoint UEs, two episodes"""
    return [n for n in _names() if n.startswith('uniform_')]


def obs_variant_names():
    """fixed-population traces of the MaxNormEnv observation (variants.py:308-332; a variable-population one rides with
    population_names() as pop_maxnorm_*)"""
    return [n for n in _names() if n.startswith('maxnorm_')]


def utility_names():
    """traces with User.util_func = 'step' (CLI --util step)"""
    return [n for n in _names() if n.startswith('utilstep_')]


def brute_names():
    """brute-force traces: the reward of every joint action per step (agent/brute_force.py, base.py:284-313)"""
    return [n for n in _names() if n.startswith('brute_')]


def population_names():
    """variable-population traces: ue_arrival / new_ue_interval on envs with max_ues > num_ue, arrays padded to max_ues"""
    return [n for n in _names() if n.startswith('pop_')]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
    cfg = json.loads(str(z['config']))
    return cfg, z


def population_kwargs(cfg):
    """extra constructor arguments of a variable-population trace (oracle and CUDA env take the same ones)"""
    arr = cfg.get('ue_arrival')
    return dict(max_ues=cfg['max_ues'], ue_arrival=None if arr is None else {int(t): int(n) for t, n in arr.items()},
                new_ue_interval=cfg.get('new_ue_interval'))


def oracle_kwargs(cfg):
    init_pos = cfg.get('init_pos')
    if init_pos is not None:
        init_pos = [tuple(p) for p in init_pos]
    return dict(kind=cfg['kind'], n_ue=cfg['n_ue'], bs_xy=[tuple(p) for p in cfg['bs_xy']], map_wh=tuple(cfg['map_wh']),
                sharing=cfg['sharing'], velocities=cfg['velocities'], seed=cfg['seed'], reward=cfg['reward'],
                episode_length=cfg['steps'], init_pos=init_pos, **{k: cfg[k] for k in ('util_func', 'obs_norm', 'obs_variant', 'obs_opts') if k in cfg},
                **({'uniform_moves': [None if u is None else tuple(u) for u in cfg['uniform_moves']]}
                   if 'uniform_moves' in cfg else {}),
                **({'sequential': True} if cfg.get('sequential') else {}))


def assert_close(a, b, what, rtol=RTOL, atol=ATOL):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    if not np.all(err <= tol):
        i = np.unravel_index(np.argmax(err - tol), err.shape) if err.shape else ()
        raise AssertionError(f"{what}: max violation at {i}: got {a[i]!r} want {b[i]!r} (err {err[i]:.3e})")


def assert_exact(a, b, what):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} mismatches, first at {bad[0].tolist()}: "
                             f"got {a[tuple(bad[0])]!r} want {b[tuple(bad[0])]!r}")


def check_against_golden(env, cfg, z, exact_floats=False, skip_float=(), episodes=None):
    """Replay the golden action sequence through `env` (reset_trace()/step()) and compare every recorded array."""
    steps, eps = cfg['steps'], cfg['episodes'] if episodes is None else episodes
    t = 0
    for ep in range(eps):
        r = env.reset_trace()
        for k in EXACT_KEYS:
            assert_exact(r[k], z['reset_' + k][ep], f'reset[{ep}].{k}')
        for k in FLOAT_KEYS:
            if k in skip_float:
                continue
            (assert_exact if exact_floats else assert_close)(r[k], z['reset_' + k][ep], f'reset[{ep}].{k}')
        for _ in range(steps):
            s = env.step(z['actions'][t])
            for k in EXACT_KEYS + EXACT_STEP_KEYS:
                assert_exact(s[k], z['step_' + k][t], f'step[{t}].{k}')
            for k in FLOAT_KEYS + FLOAT_STEP_KEYS:
                if k in skip_float:
                    continue
                (assert_exact if exact_floats else assert_close)(s[k], z['step_' + k][t], f'step[{t}].{k}')
            t += 1
