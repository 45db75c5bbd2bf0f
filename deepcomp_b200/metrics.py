"""Evaluation results of batched rollouts in the reference's result layouts.

The reference's evaluation loop (deepcomp/util/simulation.py:472-554 `Simulation.run_episode`) collects, per episode and
step, the step reward (central: the env's reward; multi-agent: the sum over the agents, simulation.py:380),
`info['scalar_metrics']` and `info['vector_metrics']`, and turns them into

  * one row per episode with mean / std of the step reward and of every scalar metric
    (simulation.py:556-588 `summarize_scalar_results`, written as CSV by `write_scalar_results`, :590-609), and
  * one table per vector metric with the columns episode, time_step, 'UE 1' .. 'UE max_ues', None where a UE is not
    there (simulation.py:611-667 `write_vector_results`, pickled data frames).

Here an "episode" is one env of a batch over one fragment: `BatchedMobileEnv.step_many(..., info=True)` /
`.rollout(..., info=True)` return reward [T, K] or [T, K, N], sum_utility [T, K], curr_dr / utility [T, K, N]; the K envs
of fragment f are episodes f*K .. f*K + K - 1.  Host-side bookkeeping over tensors the kernels already produced
(device -> host copy, then numpy / pandas): nothing here is on the step path.
"""
import numpy as np


def _host(x):
    return x.detach().cpu().numpy() if hasattr(x, 'detach') else np.asarray(x)


def step_rewards(fragment):
    """Per-step reward as the reference's loop records it: [T, K].  Multi-agent rewards are summed over the agents
    (simulation.py:380); rows of UEs that are not there are zero (padding) and do not change the sum."""
    r = _host(fragment['reward']).astype(np.float64)
    return r.sum(axis=-1) if r.ndim == 3 else r


def summarize_scalar_results(fragments, eps_duration=None):
    """
    simulation.py:556-588 over a list of fragments (dicts with 'reward' and 'sum_utility' of T steps x K envs each).

    :param eps_duration: seconds per episode, list of len(fragments) * K, or None -> a fragment's wall time is not an
        episode's: the columns are filled with NaN
    :returns: dict of column -> list, one entry per episode, in the reference's column order:
        episode, eps_duration_mean, eps_duration_std, step_reward_mean, step_reward_std, sum_utility_mean, sum_utility_std
    """
    results = {k: [] for k in ('episode', 'eps_duration_mean', 'eps_duration_std', 'step_reward_mean', 'step_reward_std',
                               'sum_utility_mean', 'sum_utility_std')}
    e = 0
    for frag in fragments:
        rew = step_rewards(frag)                                         # [T, K]
        su = _host(frag['sum_utility']).astype(np.float64)               # [T, K]
        for k in range(rew.shape[1]):
            d = float('nan') if eps_duration is None else float(eps_duration[e])
            results['episode'].append(e)
            results['eps_duration_mean'].append(d)                       # simulation.py:575-576: the value itself, twice
            results['eps_duration_std'].append(d)
            results['step_reward_mean'].append(np.mean(rew[:, k]))
            results['step_reward_std'].append(np.std(rew[:, k]))
            results['sum_utility_mean'].append(np.mean(su[:, k]))
            results['sum_utility_std'].append(np.std(su[:, k]))
            e += 1
    return results


def vector_results(fragments, max_ues=None, num_ue=None):
    """
    simulation.py:611-667: one pandas DataFrame per vector metric ('dr', 'utility'; base.py:404-409) with the columns
    episode, time_step, 'UE 1' .. 'UE max_ues'.

    :param num_ue: UEs present per step, int array [T] (or one int), for a variable population: entries of slots
        beyond it are None, as for a UE missing from the reference's metric dict.  The reference keys its dicts by UE
        id; with departures the ids of the UEs in the slots differ per env -- pass `ue_ids` columns yourself then
        (BatchedMobileEnv.ue_ids()); this helper labels slot i 'UE i+1', which is exact for a fixed population and for
        arrivals only.
    :returns: dict metric -> DataFrame; df.attrs carries 'metric' and 'num_episodes' (the reference adds its CLI / env
        metadata there, which a batch does not have)
    """
    import pandas as pd
    out = {}
    for metric, key in (('dr', 'curr_dr'), ('utility', 'utility')):
        cols = None
        rows_eps, rows_t, vals = [], [], []
        e0 = 0
        for frag in fragments:
            v = _host(frag[key]).astype(np.float64)                      # [T, K, N]
            T, K, N = v.shape
            S = N if max_ues is None else int(max_ues)
            if cols is None:
                cols = [f'UE {i + 1}' for i in range(S)]
            v = v[:, :, :S].astype(object)
            if num_ue is not None:
                n = np.broadcast_to(np.asarray(num_ue, dtype=np.int64), (T,))
                for t in range(T):
                    v[t, :, n[t]:] = None
            # the reference iterates episodes, then steps
            vals.append(np.transpose(v, (1, 0, 2)).reshape(K * T, S))
            rows_eps.append(np.repeat(np.arange(e0, e0 + K), T))
            rows_t.append(np.tile(np.arange(T), K))
            e0 += K
        data = {'episode': np.concatenate(rows_eps), 'time_step': np.concatenate(rows_t)}
        allv = np.concatenate(vals)
        for i, c in enumerate(cols):
            data[c] = list(allv[:, i])
        df = pd.DataFrame(data)
        df.attrs = {'metric': metric, 'num_episodes': e0}
        out[metric] = df
    return out


def write_results(fragments, prefix, eps_duration=None, max_ues=None, num_ue=None, metadata=None):
    """`<prefix>.csv` (scalar results, simulation.py:590-609) and `<prefix>_<metric>.pkl` (vector results, :611-667)."""
    import pandas as pd
    data = dict(metadata or {})
    data.update(summarize_scalar_results(fragments, eps_duration))
    pd.DataFrame(data=data).to_csv(f'{prefix}.csv')
    files = [f'{prefix}.csv']
    for metric, df in vector_results(fragments, max_ues, num_ue).items():
        df.attrs.update(metadata or {})
        df.to_pickle(f'{prefix}_{metric}.pkl')
        files.append(f'{prefix}_{metric}.pkl')
    return files
