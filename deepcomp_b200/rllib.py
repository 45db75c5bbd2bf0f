"""RLlib-shaped batch adapters over BatchedMobileEnv (ray==1.4.0 API, reference setup.py:14; ray is not imported).

The reference hands RLlib ONE env per rollout worker (deepcomp/util/env_setup.py:282-283) and scales by adding
worker processes (README.md:196-201).  On a B200 the natural unit is the whole batch, which RLlib consumes through its
vector interfaces:

* ``CentralVectorEnv``  -- ``ray.rllib.env.VectorEnv`` duck type (vector_reset / reset_at / vector_step /
  get_unwrapped) for the central agent: one K-env kernel launch per ``vector_step``.
* ``MultiAgentBaseEnv`` -- ``ray.rllib.env.BaseEnv`` duck type (poll / send_actions / try_reset) for the multi-agent
  env: observations / rewards / dones / infos are ``{env_id: {agent_id: ...}}`` dicts, agent ids are "1".."N"
  (env_setup.py:148-160).

Both keep obs as numpy views of ONE host copy of the batch per step; per-env Python dicts are only built at this boundary
(for host-buffer stepping without dicts use ``BatchedMobileEnv.step_host`` / ``step_many_host`` with pinned buffers).
Episode ends follow the reference: ``done`` is never set by the env (base.py:371-381); RLlib's ``horizon`` calls
``reset_at`` / ``try_reset``.
"""
import numpy as np
import torch

from . import spaces
from .batched import BatchedMobileEnv


class _BatchAdapter:
    kind = None

    def __init__(self, num_envs, **scenario):
        scenario = dict(scenario)
        scenario['kind'] = self.kind
        self.batch = BatchedMobileEnv(num_envs=num_envs, **scenario)
        # variable UE population (base.py:429-443, 592-617): the batch steps in lockstep, but WHO leaves is drawn per env,
        # so the agent ids per slot differ between envs -- they are read back from the device after every event
        self.dynamic = self.batch._dynamic
        self.num_envs = num_envs
        self.n_ue, self.n_bs = self.batch.n_ue, self.batch.n_bs
        self.agent_ids = [str(i + 1) for i in range(self.n_ue)]
        self._ids = None            # dynamic: int array [K, active_ues] of User.id per slot
        self._obs = None
        # 'dr' is Box(0, 1) for the RelNorm observation, Box(-1, 1) for MaxNorm (variants.py:259, 313-317)
        self._dr_low = -1 if self.batch.obs_norm == 'max' else 0

    def _central_obs(self, flat):
        nm = self.n_ue * self.n_bs
        return {'connected': flat[:nm].astype(np.int8), 'dr': flat[nm:2 * nm], 'utility': flat[2 * nm:]}

    def _agent_obs(self, row):
        m = self.n_bs
        return {'connected': row[:m].astype(np.int8), 'dr': row[m:2 * m], 'ues_at_bs': row[2 * m:3 * m],
                'util_at_bs': row[3 * m:4 * m], 'utility': row[4 * m:]}

    def _refresh_ids(self):
        if self.dynamic:
            self._ids = self.batch.ue_ids()

    def _agents_of(self, k):
        """agent ids (multi_agent.py:21-37: the UE ids as strings) of the UEs present in env k, in slot order"""
        if not self.dynamic:
            return self.agent_ids
        return [str(int(i)) for i in self._ids[k]]

    def _info(self, k, info):
        return {'time': int(self._time[k]),
                'scalar_metrics': {'sum_utility': float(info['sum_utility'][k])}}

    def close(self):
        self.batch.close()


class CentralVectorEnv(_BatchAdapter):
    kind = 'central'

    def __init__(self, num_envs, **scenario):
        super().__init__(num_envs, **scenario)
        n, m = self.n_ue, self.n_bs
        self.action_space = spaces.MultiDiscrete([m + 1] * n)
        self.observation_space = spaces.Dict({
            'connected': spaces.MultiBinary(n * m), 'dr': spaces.Box(low=self._dr_low, high=1, shape=(n * m,)),
            'utility': spaces.Box(low=-1, high=1, shape=(n,))})
        self._time = np.zeros(num_envs, dtype=np.int64)

    def vector_reset(self):
        obs = self.batch.reset().cpu().numpy()
        self._time[:] = 0
        self._obs = obs
        return [self._central_obs(obs[k]) for k in range(self.num_envs)]

    def reset_at(self, index=0):
        if self.dynamic:
            raise NotImplementedError("a batch with a variable UE population resets as a whole (vector_reset)")
        obs = self.batch.reset(env_ids=[index]).cpu().numpy()
        self._time[index] = 0
        self._obs = obs
        return self._central_obs(obs[index])

    def vector_step(self, actions):
        a = torch.as_tensor(np.asarray(actions, dtype=np.int32).reshape(self.num_envs, self.n_ue),
                            device=self.batch.device)
        obs, rew, _, info = self.batch.step(a)
        self.batch.check_errors()
        obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
        info = {k: v.cpu().numpy() for k, v in info.items()}
        self._time += 1
        self._obs = obs
        K = self.num_envs
        return ([self._central_obs(obs[k]) for k in range(K)], [float(rew[k]) for k in range(K)], [None] * K,
                [self._info(k, info) for k in range(K)])

    def get_unwrapped(self):
        return []


class MultiAgentBaseEnv(_BatchAdapter):
    kind = 'multi'

    def __init__(self, num_envs, **scenario):
        super().__init__(num_envs, **scenario)
        m = self.n_bs
        self.action_space = spaces.Discrete(m + 1)
        self.observation_space = spaces.Dict({
            'connected': spaces.MultiBinary(m), 'dr': spaces.Box(low=self._dr_low, high=1, shape=(m,)),
            'utility': spaces.Box(low=-1, high=1, shape=(1,)), 'ues_at_bs': spaces.Box(low=0, high=1, shape=(m,)),
            'util_at_bs': spaces.Box(low=-1, high=1, shape=(m,))})
        self._time = np.zeros(num_envs, dtype=np.int64)
        self._pending = None
        obs = self.batch.reset().cpu().numpy()
        self._refresh_ids()
        self._pending = (self._obs_dicts(obs), {k: {} for k in range(num_envs)},
                         {k: {'__all__': None} for k in range(num_envs)}, {k: {} for k in range(num_envs)})

    def _obs_dicts(self, obs, only=None):
        ks = range(self.num_envs) if only is None else only
        return {k: {aid: self._agent_obs(obs[k, i]) for i, aid in enumerate(self._agents_of(k))} for k in ks}

    def poll(self):
        """-> (obs, rewards, dones, infos, off_policy_actions), each {env_id: {agent_id: value}}"""
        obs, rew, dones, infos = self._pending
        self._pending = ({}, {}, {}, {})
        return obs, rew, dones, infos, {}

    def send_actions(self, action_dict):
        """action_dict: {env_id: {agent_id: action}}; agents / envs missing from the dict do nothing
        (multi_agent.py:21-30)"""
        a = np.zeros((self.num_envs, self.n_ue), dtype=np.int32)
        for k, acts in action_dict.items():
            if self.dynamic:
                slot = {int(i): j for j, i in enumerate(self._ids[k])}
                for aid, v in acts.items():
                    if int(aid) in slot:                 # a UE that is not in the list does nothing (multi_agent.py:21-30)
                        a[k, slot[int(aid)]] = int(v)
            else:
                for aid, v in acts.items():
                    a[k, int(aid) - 1] = int(v)
        obs, rew, _, info = self.batch.step(torch.as_tensor(a, device=self.batch.device))
        self._refresh_ids()                              # arrivals / departures of this step are in the returned dicts
        self.batch.check_errors()
        obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
        info = {k: v.cpu().numpy() for k, v in info.items()}
        self._time += 1
        K = self.num_envs
        rewards = {k: {aid: float(rew[k, i]) for i, aid in enumerate(self._agents_of(k))} for k in range(K)}
        dones = {k: dict({aid: None for aid in self._agents_of(k)}, __all__=None) for k in range(K)}
        infos = {k: {aid: self._info(k, info) for aid in self._agents_of(k)} for k in range(K)}
        self._pending = (self._obs_dicts(obs), rewards, dones, infos)

    def try_reset(self, env_id=None):
        if self.dynamic and env_id is not None:
            raise NotImplementedError("a batch with a variable UE population resets as a whole (try_reset(None))")
        ids = None if env_id is None else [env_id]
        obs = self.batch.reset(env_ids=ids).cpu().numpy()
        self._refresh_ids()
        if env_id is None:
            self._time[:] = 0
            return self._obs_dicts(obs)
        self._time[env_id] = 0
        return self._obs_dicts(obs, only=[env_id])[env_id]

    def get_unwrapped(self):
        return []

    def stop(self):
        self.close()
