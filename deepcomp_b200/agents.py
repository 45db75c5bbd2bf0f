"""Baseline agents of the reference (deepcomp/agent/heuristics.py, deepcomp/agent/dummy.py) in two forms.

* Host / tensor form: the same classes and constructor arguments as the reference; the decision rules are written over
  whole observation batches (``compute_actions(obs [..., N, 4M+1]) -> int32 [..., N]``, torch ops on whichever device
  the observations live on), and the reference's per-UE ``compute_action(obs_dict, policy_id)`` is a thin wrapper -- so
  the reference's evaluation loop (``Simulation.apply_action_multi_agent``, deepcomp/util/simulation.py:351-380) runs
  unchanged on the facades of ``deepcomp_b200.env``.
* Device form: ``agent.device_policy(batch)`` describes the same decision rule to the CUDA step kernel, which then
  drives all K envs for a whole fragment without a host round trip (``BatchedMobileEnv.rollout``): the physics warps
  evaluate the rule from the state they already hold ("highest dr" = "smallest distance", "dr >= eps * best" as a
  squared-distance ratio), one step ahead of the observation the host loop would have needed.
"""
import random

import numpy as np
import torch

POLICY_KIND = {'3gpp': 1, 'fullcomp': 2, 'dynamic': 3, 'static': 4, 'fixed': 5, 'random': 6}


class MultiAgent:
    """deepcomp/agent/base.py:13-19"""

    def __init__(self):
        self.central_agent = False

    def compute_action(self, observation, policy_id):
        raise NotImplementedError("This needs to be implemented in the child class")


class CentralAgent:
    """deepcomp/agent/base.py:4-10"""

    def __init__(self):
        self.central_agent = True

    def compute_action(self, observation):
        raise NotImplementedError("This needs to be implemented in the child class")


def _split(obs, n_bs=None):
    """(connected, dr) as bool / float tensors [..., M] from a packed multi-agent observation [..., 4M+1] (the layout of
    BatchedMobileEnv: connected | dr | ues_at_bs | util_at_bs | utility) or from one of the facades' obs dicts"""
    if isinstance(obs, dict):
        return torch.as_tensor(np.asarray(obs['connected'])) > 0, torch.as_tensor(np.asarray(obs['dr'], dtype=np.float64))
    obs = torch.as_tensor(obs)
    m = (obs.shape[-1] - 1) // 4 if n_bs is None else n_bs
    return obs[..., :m] > 0, obs[..., m:2 * m]


def _first_true(mask):
    """index of the first True along the last axis (0 where there is none)"""
    return torch.argmax(mask.to(torch.int8), dim=-1)


def _argmax_first(x):
    """np.argmax semantics (first maximum) along the last axis"""
    return _first_true(x == x.max(dim=-1, keepdim=True).values)


def _pick_in_set(conn, dr, selected):
    """Common tail of DynamicSelection / StaticClustering for a whole batch (heuristics.py:93-108, 177-187): leave the
    first connected BS outside the set; else join the strongest BS of the set that is not connected yet; else no-op."""
    leave = conn & ~selected
    want = selected & ~conn
    strongest = _argmax_first(torch.where(want, dr, torch.full_like(dr, -float('inf'))))
    act = torch.where(want.any(dim=-1), strongest + 1, torch.zeros_like(strongest))
    return torch.where(leave.any(dim=-1), _first_true(leave) + 1, act).to(torch.int32)


class _BatchedHeuristic(MultiAgent):
    """
    The per-UE decision rules evaluated for ALL UEs of ALL envs at once: ``compute_actions(obs)`` takes the packed
    observation tensor of BatchedMobileEnv ([..., N, 4M+1], host or device) and returns int32 actions [..., N];
    ``compute_action(obs, policy_id)`` is the reference's per-UE signature (one obs dict) on top of it.
    """

    def compute_actions(self, obs, n_bs=None):
        raise NotImplementedError

    def compute_action(self, obs, policy_id=None):
        return int(self.compute_actions(obs))


class Heuristic3GPP(_BatchedHeuristic):
    """At most one BS, the one with the highest SNR; leave any other BS first (heuristics.py:13-38)."""

    def compute_actions(self, obs, n_bs=None):
        conn, dr = _split(obs, n_bs)
        best = _argmax_first(dr)
        at_best = torch.gather(conn, -1, best.unsqueeze(-1)).squeeze(-1)
        act = torch.where(conn.any(dim=-1), _first_true(conn) + 1, best + 1)
        return torch.where(at_best, torch.zeros_like(act), act).to(torch.int32)

    def device_policy(self, batch=None):
        return dict(kind='3gpp')


class FullCoMP(_BatchedHeuristic):
    """Connect to every BS, strongest first (heuristics.py:41-65)."""

    def compute_actions(self, obs, n_bs=None):
        conn, dr = _split(obs, n_bs)
        return _pick_in_set(conn, dr, torch.ones_like(conn))

    def device_policy(self, batch=None):
        return dict(kind='fullcomp')


class DynamicSelection(_BatchedHeuristic):
    """Strongest BS and all BS within epsilon * SNR of it (heuristics.py:68-108)."""

    def __init__(self, epsilon):
        super().__init__()
        assert 0 <= epsilon <= 1, f"Scaling factor epsilon must be within [0,1] but is {epsilon}."   # cli.py:100
        self.epsilon = epsilon

    def compute_actions(self, obs, n_bs=None):
        conn, dr = _split(obs, n_bs)
        return _pick_in_set(conn, dr, dr >= dr.max(dim=-1, keepdim=True).values * self.epsilon)

    def device_policy(self, batch=None):
        return dict(kind='dynamic', epsilon=float(self.epsilon))


class StaticClustering(_BatchedHeuristic):
    """Static, non-overlapping clusters of `cluster_size` closest cells (heuristics.py:111-187)."""

    def __init__(self, cluster_size, bs_list, seed=None, clusters=None):
        super().__init__()
        self.cluster_size, self.bs_list, self.seed = cluster_size, list(bs_list), seed
        self.rng = random.Random()
        self.rng.seed(seed)
        self.clusters = clusters if clusters is not None else self.build_clusters()

    def build_clusters(self):
        """heuristics.py:132-167; returns {bs index: set of bs indices in the same cluster}"""
        clusters = {}
        remaining = list(range(len(self.bs_list)))
        curr = []
        while len(remaining) > 0:
            if len(curr) == 0:
                bs = self.rng.choice(remaining)
                curr.append(bs)
                remaining.remove(bs)
            else:
                cx = np.mean([self.bs_list[b].pos.x for b in curr])
                cy = np.mean([self.bs_list[b].pos.y for b in curr])
                closest = min(remaining, key=lambda b: np.sqrt((cx - self.bs_list[b].pos.x) ** 2
                                                                + (cy - self.bs_list[b].pos.y) ** 2))
                curr.append(closest)
                remaining.remove(closest)
            if len(curr) == self.cluster_size:
                for b in curr:
                    clusters[b] = set(curr)
                curr = []
        for b in curr:
            clusters[b] = set(curr)
        return clusters

    def compute_actions(self, obs, n_bs=None):
        conn, dr = _split(obs, n_bs)
        m = conn.shape[-1]
        member = torch.zeros((m, m), dtype=torch.bool, device=conn.device)      # member[b, c]: c is in the cluster of b
        for b, cluster in self.clusters.items():
            member[b, sorted(cluster)] = True
        return _pick_in_set(conn, dr, member[_argmax_first(dr)])

    def cluster_masks(self):
        m = np.zeros(len(self.bs_list), dtype=np.uint64)
        for b, members in self.clusters.items():
            for c in members:
                m[b] |= np.uint64(1) << np.uint64(c)
        return m

    def device_policy(self, batch=None):
        return dict(kind='static', cluster_masks=self.cluster_masks())


class RandomAgent(CentralAgent):
    """dummy.py:6-22.  The host form samples the action space; the device form uses its own counter-based RNG."""

    def __init__(self, action_space, num_vec_envs=None, seed=None):
        super().__init__()
        self.action_space, self.num_vec_envs, self.seed = action_space, num_vec_envs, seed
        self.action_space.seed(seed)

    def compute_action(self, observation):
        if self.num_vec_envs is None:
            return self.action_space.sample()
        return [self.action_space.sample() for _ in range(self.num_vec_envs)]

    def device_policy(self, batch=None):
        return dict(kind='random', seed=0 if self.seed is None else int(self.seed))


class FixedAgent(CentralAgent):
    """Always the same action, `noop_interval` no-op steps in between (dummy.py:25-50)."""

    def __init__(self, action, noop_interval=0, num_vec_envs=None):
        super().__init__()
        self.action, self.noop_interval, self.num_vec_envs = action, noop_interval, num_vec_envs
        self.noop_counter = noop_interval
        self.device_calls = 0          # steps this agent has driven on the device (phase of its no-op interval)

    def compute_action(self, observation):
        if self.noop_counter < self.noop_interval:
            action = np.zeros(len(self.action))
            self.noop_counter += 1
        else:
            action = self.action
            self.noop_counter = 0
        if self.num_vec_envs is None:
            return action
        return [action for _ in range(self.num_vec_envs)]

    def device_policy(self, batch=None):
        return dict(kind='fixed', fixed_action=np.asarray(self.action, dtype=np.int32),
                    noop_interval=int(self.noop_interval), calls_before=int(self.device_calls))


class BruteForceAgent(CentralAgent):
    """
    Reference deepcomp/agent/brute_force.py:10-94: tests every joint action on the env and takes the best one (first
    maximum).  The reference walks the (M + 1)^N candidates one by one through MobileEnv.test_ue_actions
    (base.py:284-313); here they are evaluated in one launch (`dcb_test_actions`), so `num_workers` has no meaning.
    `env` is a CentralRelNormEnv facade (its K = 1 batch is used) or a BatchedMobileEnv (+ `env_index`).
    """

    def __init__(self, num_workers=1, env=None, env_index=0):
        super().__init__()
        self.num_workers = num_workers
        self.env = env
        self.env_index = env_index

    def _batch(self):
        assert self.env is not None, "Set agent's env before computing actions."       # brute_force.py:81
        return getattr(self.env, '_batch', self.env)

    def get_ith_action(self, i):
        """brute_force.py:59-62"""
        return self._batch().candidate_action(i)

    def compute_action(self, observation):
        """brute_force.py:79-94"""
        action, _ = self._batch().best_joint_action(self.env_index)
        return action
