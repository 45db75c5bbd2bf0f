"""Baseline agents of the reference (deepcomp/agent/heuristics.py, deepcomp/agent/dummy.py) in two forms.

* Host form: the same classes, constructor arguments and ``compute_action`` signatures as the reference, operating on
  the obs dicts the facades in ``deepcomp_b200.env`` return -- so the reference's evaluation loop
  (``Simulation.apply_action_multi_agent``, deepcomp/util/simulation.py:351-380) runs unchanged.
* Device form: ``agent.device_policy(batch)`` describes the same decision rule to the CUDA step kernel, which then
  drives all K envs for a whole fragment without a host round trip (``BatchedMobileEnv.rollout``): the physics warps
  evaluate the rule from the state they already hold ("highest dr" = "smallest distance", "dr >= eps * best" as a
  squared-distance ratio), one step ahead of the observation the host loop would have needed.
"""
import random

import numpy as np

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


class Heuristic3GPP(MultiAgent):
    """Always at most one BS: the one with the highest SNR (heuristics.py:13-38)."""

    def compute_action(self, obs, policy_id=None):
        best_bs = int(np.argmax(obs['dr']))
        if obs['connected'][best_bs]:
            return 0
        if sum(obs['connected']) > 0:
            return list(obs['connected']).index(1) + 1
        return best_bs + 1

    def device_policy(self, batch=None):
        return dict(kind='3gpp')


class FullCoMP(MultiAgent):
    """Greedily connect to all BS, strongest first (heuristics.py:41-65)."""

    def compute_action(self, obs, policy_id=None):
        disconn_bs = [idx for idx, conn in enumerate(obs['connected']) if not conn]
        if len(disconn_bs) == 0:
            return 0
        best_bs = disconn_bs[0]
        best_dr = obs['dr'][best_bs]
        for bs in disconn_bs:
            if obs['dr'][bs] > best_dr:
                best_bs, best_dr = bs, obs['dr'][bs]
        return best_bs + 1

    def device_policy(self, batch=None):
        return dict(kind='fullcomp')


def _select_within(obs, selected):
    """Common tail of DynamicSelection / StaticClustering (heuristics.py:93-108, 177-187)."""
    connected = [idx for idx, conn in enumerate(obs['connected']) if conn]
    for bs in connected:
        if bs not in selected:
            return bs + 1
    for bs in sorted(selected, key=lambda idx: obs['dr'][idx], reverse=True):
        if not obs['connected'][bs]:
            return bs + 1
    return 0


class DynamicSelection(MultiAgent):
    """Strongest BS and all BS within epsilon * SNR of it (heuristics.py:68-108)."""

    def __init__(self, epsilon):
        super().__init__()
        assert 0 <= epsilon <= 1, f"Scaling factor epsilon must be within [0,1] but is {epsilon}."   # cli.py:100
        self.epsilon = epsilon

    def compute_action(self, obs, policy_id=None):
        threshold = max(obs['dr']) * self.epsilon
        return _select_within(obs, [idx for idx, snr in enumerate(obs['dr']) if snr >= threshold])

    def device_policy(self, batch=None):
        return dict(kind='dynamic', epsilon=float(self.epsilon))


class StaticClustering(MultiAgent):
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

    def compute_action(self, obs, policy_id=None):
        return _select_within(obs, sorted(self.clusters[int(np.argmax(obs['dr']))]))

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
                    noop_interval=int(self.noop_interval))


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
