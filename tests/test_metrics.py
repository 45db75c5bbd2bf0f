"""CPU: deepcomp_b200.metrics against the reference's own result code (deepcomp/util/simulation.py:556-667).

`simulation.py` cannot be imported here (it needs RLlib), so the two functions are cut out of its source with `ast` and
executed unmodified -- test infrastructure only, this container only (skipped where /root/reference is absent)."""
import ast
import os
import types
from collections import defaultdict

import numpy as np
import pytest

from deepcomp_b200 import metrics

REF_SIM = os.path.join(os.environ.get('DEEPCOMP_REFERENCE', '/root/reference'), 'deepcomp', 'util', 'simulation.py')


def _fragments(seed, n_frag=2, T=7, K=3, N=4, multi=True):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_frag):
        out.append(dict(reward=rng.normal(size=(T, K, N) if multi else (T, K)).astype(np.float32),
                        sum_utility=rng.normal(size=(T, K)).astype(np.float32) * 20,
                        curr_dr=rng.random((T, K, N)).astype(np.float32) * 50,
                        utility=rng.normal(size=(T, K, N)).astype(np.float32) * 10))
    return out


def _reference_inputs(frags, num_ue=None):
    """what Simulation.run_episode collects (simulation.py:472-554), one entry per (fragment, env)"""
    rewards, scalar, vector = [], [], []
    for f in frags:
        T, K, N = f['curr_dr'].shape
        r = f['reward'].astype(np.float64)
        for k in range(K):
            rewards.append([float(r[t, k].sum()) if r.ndim == 3 else float(r[t, k]) for t in range(T)])   # :380
            scalar.append([{'sum_utility': float(f['sum_utility'][t, k])} for t in range(T)])              # base.py:402
            ep = []
            for t in range(T):
                n = N if num_ue is None else int(num_ue[t])
                ep.append({'dr': {f'UE {i + 1}': float(f['curr_dr'][t, k, i]) for i in range(n)},          # base.py:404-409
                           'utility': {f'UE {i + 1}': float(f['utility'][t, k, i]) for i in range(n)}})
            vector.append(ep)
    return rewards, scalar, vector


def _reference_functions():
    tree = ast.parse(open(REF_SIM).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'Simulation')
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ('summarize_scalar_results',
                                                                                 'write_vector_results')]
    for f in fns:
        f.decorator_list = []
    mod = ast.Module(body=fns, type_ignores=[])
    import pandas as pd
    ns = {'np': np, 'pd': pd, 'defaultdict': defaultdict}
    exec(compile(mod, REF_SIM, 'exec'), ns)
    return ns['summarize_scalar_results'], ns['write_vector_results']


@pytest.mark.skipif(not os.path.exists(REF_SIM), reason='/root/reference not present (GPU box)')
@pytest.mark.parametrize('multi', [True, False])
def test_scalar_results_match_reference_summary(multi):
    frags = _fragments(1, multi=multi)
    dur = list(np.arange(6) * 0.5 + 1)
    rewards, scalar, _ = _reference_inputs(frags)
    want = _reference_functions()[0](dur, rewards, scalar)
    got = metrics.summarize_scalar_results(frags, eps_duration=dur)
    assert list(got.keys()) == list(want.keys())
    for k in want:
        np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64), np.asarray(want[k], dtype=np.float64),
                                   rtol=1e-12, atol=1e-12, err_msg=k)


@pytest.mark.skipif(not os.path.exists(REF_SIM), reason='/root/reference not present (GPU box)')
@pytest.mark.parametrize('num_ue', [None, [2, 2, 3, 4, 4, 4, 3]])
def test_vector_results_match_reference_frames(num_ue, tmp_path):
    frags = _fragments(2)
    _, _, vector = _reference_inputs(frags, num_ue)
    fake = types.SimpleNamespace(
        env=types.SimpleNamespace(max_ues=4), metadata={}, cli_args=types.SimpleNamespace(),
        env_config={'ue_arrival': None, 'map': 'm', 'ue_list': [], 'bs_list': []}, test_dir=str(tmp_path),
        result_filename='res', log=types.SimpleNamespace(info=lambda *a, **k: None))
    want = _reference_functions()[1](fake, vector)
    got = metrics.vector_results(frags, max_ues=4, num_ue=num_ue)
    assert [df.attrs['metric'] for df in want] == list(got.keys())
    for df in want:
        g = got[df.attrs['metric']]
        assert list(g.columns) == list(df.columns) and len(g) == len(df)
        assert g.attrs['num_episodes'] == df.attrs['num_episodes']
        for c in df.columns:
            a, b = g[c].to_numpy(dtype=np.float64, na_value=np.nan), df[c].to_numpy(dtype=np.float64, na_value=np.nan)
            np.testing.assert_array_equal(a, b, err_msg=c)


def test_write_results_files_and_layout(tmp_path):
    import pandas as pd
    frags = _fragments(3, n_frag=1, T=5, K=2, N=3)
    files = metrics.write_results(frags, str(tmp_path / 'run'), max_ues=3, metadata={'alg': 'ppo'})
    assert [os.path.basename(f) for f in files] == ['run.csv', 'run_dr.pkl', 'run_utility.pkl']
    csv = pd.read_csv(files[0], index_col=0)
    assert list(csv.columns) == ['alg', 'episode', 'eps_duration_mean', 'eps_duration_std', 'step_reward_mean',
                                 'step_reward_std', 'sum_utility_mean', 'sum_utility_std'] and len(csv) == 2
    np.testing.assert_allclose(csv['step_reward_mean'][1], frags[0]['reward'].astype(np.float64).sum(-1)[:, 1].mean())
    df = pd.read_pickle(files[1])
    assert list(df.columns) == ['episode', 'time_step', 'UE 1', 'UE 2', 'UE 3'] and len(df) == 10
    assert df['episode'].tolist() == [0] * 5 + [1] * 5 and df['time_step'].tolist() == list(range(5)) * 2
    np.testing.assert_allclose(df['UE 2'].to_numpy(dtype=np.float64)[5:], frags[0]['curr_dr'][:, 1, 1])
