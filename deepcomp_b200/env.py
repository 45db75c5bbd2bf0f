"""Drop-in env classes with the reference's gym.Env / RLlib MultiAgentEnv surface, backed by the CUDA batch.

``CentralRelNormEnv`` and ``MultiAgentMobileEnv`` mirror the classes ``get_env_class`` selects in the reference
(deepcomp/util/env_setup.py:23-37 -> deepcomp/env/multi_ue/central.py:143-152, deepcomp/env/multi_ue/multi_agent.py:6-107):
same constructor (``cls(env_config)`` with the dict of env_setup.py:247-256), same ``reset() -> obs``,
``step(action) -> (obs, reward, done, info)``, ``observation_space`` / ``action_space``, ``done`` is ``None``
(base.py:371-381), ``info`` as in base.py:383-411.  ``env_config`` may carry the reference's own Map / Basestation /
User objects or the descriptors of ``deepcomp_b200.entities`` -- only attributes are read (duck typing).

One instance = one env (K = 1 slice of the batched kernel), which is what RLlib's rollout workers construct; for
throughput use ``BatchedMobileEnv`` / ``deepcomp_b200.rllib`` directly.  Extra, optional ``env_config`` keys:
``device`` (CUDA device, default current), ``log_metrics`` False skips the info metrics (base.py:389-390).
"""
import numpy as np
import torch

from . import spaces
from .batched import BatchedMobileEnv
from .entities import Point


def _parse_env_config(env_config):
    """Read the scenario out of a reference-style env_config (base.py:27-84)."""
    required = ['episode_length', 'map', 'bs_list', 'ue_list', 'new_ue_interval', 'ue_arrival', 'seed', 'rand_episodes',
                'log_metrics', 'max_ues']
    for k in required:
        if k not in env_config:
            raise KeyError(f"env_config is missing {k!r} (reference MobileEnv.__init__, base.py:27-84)")
    m, bs_list, ue_list = env_config['map'], env_config['bs_list'], env_config['ue_list']
    max_ues = env_config['max_ues']
    if max_ues is None:
        # base.py:191-209 get_max_num_ue
        max_ues = len(ue_list)
        if env_config['new_ue_interval'] is not None:
            # "eps_length - 1 because time is increased before checking done and t=eps_length is never reached"
            max_ues += int((env_config['episode_length'] - 1) / env_config['new_ue_interval'])
        if env_config['ue_arrival'] is not None:
            curr = max_ues                      # running number of UEs over the arrival / departure sequence
            for arrival in env_config['ue_arrival'].values():
                curr += arrival
                max_ues = max(max_ues, curr)
    assert max_ues >= len(ue_list)                                                        # base.py:84
    velocities, init_pos, pauses, borders, utils, uniform = [], [], set(), set(), set(), []
    for ue in ue_list:
        if getattr(ue, 'util_func', 'log') not in ('log', 'step'):
            raise NotImplementedError(f"Utility function {ue.util_func} not implemented!")   # user.py:92
        utils.add((getattr(ue, 'util_func', 'log'), getattr(ue, 'dr_req', 1)))
        mv = ue.movement
        init_pos.append((ue.init_pos_x, ue.init_pos_y))
        if hasattr(mv, 'init_move_x'):                       # UniformMovement (movement.py:26-80)
            uniform.append((mv.init_move_x, mv.init_move_y))
            velocities.append(0)
            continue
        if not hasattr(mv, 'init_velocity'):
            raise NotImplementedError("movement must be a RandomWaypoint or a UniformMovement (movement.py:26-181)")
        uniform.append(None)
        velocities.append(mv.init_velocity)
        pauses.add(mv.pause_duration)
        borders.add(mv.border_buffer)
    if not pauses:
        pauses, borders = {2}, {10}                          # only UniformMovement UEs: the RandomWaypoint defaults
    if len(pauses) != 1 or len(borders) != 1:
        raise NotImplementedError("per-UE pause_duration / border_buffer are not supported")
    if len(utils) != 1:
        raise NotImplementedError("per-UE utility functions / required rates are not supported")
    util_func, dr_req = utils.pop()
    return dict(n_ue=len(ue_list), bs_xy=[(float(bs.pos.x), float(bs.pos.y)) for bs in bs_list],
                map_wh=(int(m.width), int(m.height)), sharing=[bs.sharing_model for bs in bs_list],
                velocities=velocities, init_pos=init_pos, pause_duration=pauses.pop(), border_buffer=borders.pop(),
                episode_length=env_config['episode_length'], rand_episodes=bool(env_config['rand_episodes']),
                max_ues=int(max_ues), ue_arrival=env_config['ue_arrival'], new_ue_interval=env_config['new_ue_interval'],
                util_func=util_func, dr_req=dr_req,
                uniform_moves=uniform if any(u is not None for u in uniform) else None)


class _MobileEnvFacade:
    """Common part of the two facades: the attributes and helpers of MobileEnv (base.py:20-143, 383-411)."""
    metadata = {'render.modes': ['human']}
    _kind = None
    _obs_norm = 'rel'           # 'dr' normalisation: RelNormEnv (variants.py:276-284) / 'max' = MaxNormEnv (:308-332)
    _obs_variant = None         # 'normdr' / 'datarate': the data-rate observation classes (variants.py:42-250)

    def _obs_options(self):
        return None

    def __init__(self, env_config):
        self.env_config = env_config
        sc = _parse_env_config(env_config)
        self.episode_length = sc['episode_length']
        self.map, self.bs_list, self.ue_list = env_config['map'], env_config['bs_list'], env_config['ue_list']
        self.original_ue_list = list(self.ue_list)
        self.new_ue_interval, self.ue_arrival = sc['new_ue_interval'], sc['ue_arrival']
        self.env_seed = env_config['seed']
        self.rand_episodes = sc['rand_episodes']
        self.log_metrics = env_config['log_metrics']
        self.dashboard = env_config.get('dashboard', False)
        self.ue_details = env_config.get('ue_details', False)
        self.max_ues = sc['max_ues']
        self._ue_by_id = {int(ue.id): ue for ue in self.ue_list}
        self.reward_agg = env_config['reward']                      # central.py:19 / multi_agent.py:19
        self.time = 0
        self.total_utility = 0
        self.obs = None
        seed = self.env_seed
        if seed is None:
            # the reference falls back to OS entropy (random.Random() unseeded); draw one seed the same way
            seed = int(np.random.SeedSequence().generate_state(1)[0])
        self._scenario = sc
        self._device = env_config.get('device', None)
        self._make_batch(seed)
        m = self.num_bs
        # actions: select a BS to connect to / disconnect from, or noop (variants.py:17)
        self._ue_action_space = spaces.Discrete(m + 1)
        # variants.py:255-269
        self.obs_space_dict = {
            'connected': spaces.MultiBinary(m),
            'dr': spaces.Box(low=0, high=1, shape=(m,)),
            'utility': spaces.Box(low=-1, high=1, shape=(1,)),
            'ues_at_bs': spaces.Box(low=0, high=1, shape=(m,)),
            'util_at_bs': spaces.Box(low=-1, high=1, shape=(m,)),
        }

    def _make_batch(self, seed):
        sc = self._scenario
        self._batch = BatchedMobileEnv(
            num_envs=1, n_ue=sc['n_ue'], bs_xy=sc['bs_xy'], map_wh=sc['map_wh'], kind=self._kind,
            sharing=sc['sharing'], velocities=sc['velocities'], seeds=[seed], reward=self.reward_agg,
            episode_length=sc['episode_length'], rand_episodes=sc['rand_episodes'], init_pos=sc['init_pos'],
            pause_duration=sc['pause_duration'], border_buffer=sc['border_buffer'], device=self._device,
            max_ues=sc['max_ues'], ue_arrival=sc['ue_arrival'], new_ue_interval=sc['new_ue_interval'],
            util_func=sc['util_func'], dr_req=sc['dr_req'], obs_norm=self._obs_norm, uniform_moves=sc['uniform_moves'],
            obs_variant=self._obs_variant, obs_opts=self._obs_options(), interference=bool(self.env_config.get('interference', False)))

    # ---- MobileEnv attributes
    @property
    def num_bs(self):
        return len(self.bs_list)

    @property
    def num_ue(self):
        return len(self.ue_list)

    def get_max_num_ue(self):
        """base.py:191-209: the maximum number of UEs within an episode"""
        max_ues = self.num_ue
        if self.new_ue_interval is not None:
            # "eps_length - 1 because time is increased before checking done and t=eps_length is never reached"
            max_ues = self.num_ue + int((self.episode_length - 1) / self.new_ue_interval)
        if self.ue_arrival is not None:
            curr_ues = max_ues
            for arrival in self.ue_arrival.values():
                curr_ues += arrival
                if curr_ues > max_ues:
                    max_ues = curr_ues
        return max_ues

    def get_num_diff_ues(self):
        """base.py:211-225: the number of DIFFERENT UEs over an episode (env_setup.py:295 makes one policy per UE of it)"""
        max_ues = self.get_max_num_ue()
        if self.ue_arrival is None:
            return max_ues
        num_diff_ues = self.num_ue
        for arrival in self.ue_arrival.values():
            if arrival > 0:
                num_diff_ues += arrival
        return num_diff_ues

    def seed(self, seed=None):
        """
        base.py:132-143: re-seed every UE's RNGs.  As in the reference this only matters with rand_episodes=True:
        otherwise the next reset() re-applies env_config['seed'] (base.py:171-173) and the manual seed is forgotten.
        """
        if seed is not None and self.rand_episodes:
            self._batch.close()
            self._make_batch(seed)

    @staticmethod
    def set_log_level(log_dict):
        """base.py:145-152 (the B200 env does not log from the hot path)"""

    def done(self):
        """base.py:371-381: no natural episode end; RLlib's horizon resets the env"""
        return None

    def close(self):
        self._batch.close()

    def render(self, *a, **k):
        raise NotImplementedError("rendering (base.py:468-590) is outside the B200 hot path")

    # ---- state views for callers that read the entity objects (simulation.py:472-554, callbacks.py:16-33)
    def _sync_entities(self, curr_dr, utility):
        st = self._batch.get_state()
        mask = self._batch.mask_matrix(st['mask'])[0]
        for i, ue in enumerate(self.ue_list):
            try:
                ue.pos = Point(float(st['pos'][0, i, 0]), float(st['pos'][0, i, 1]))
                ue.curr_dr, ue.utility = float(curr_dr[i]), float(utility[i])
                ue.ewma_dr = float(st['ewma'][0, i])
            except AttributeError:
                pass    # reference User objects expose these as read-only properties
        for b, bs in enumerate(self.bs_list):
            try:
                bs.num_conn_ues = int(mask[:, b].sum())
            except AttributeError:
                pass

    def _info(self, curr_dr, utility, sum_utility):
        """base.py:383-411"""
        if not self.log_metrics:
            return {'time': self.time}
        return {
            'time': self.time,
            'scalar_metrics': {'sum_utility': float(sum_utility)},
            'vector_metrics': {
                'dr': {f'UE {ue}': float(curr_dr[i]) for i, ue in enumerate(self.ue_list)},
                'utility': {f'UE {ue}': float(utility[i]) for i, ue in enumerate(self.ue_list)},
            },
        }

    def _device_actions(self, per_ue):
        a = np.zeros((1, self.max_ues), dtype=np.int32)
        a[0, :len(per_ue)] = per_ue
        return torch.as_tensor(a, device=self._batch.device)

    def _sync_ue_list(self):
        """arrivals / departures (base.py:592-617): rebuild ue_list from the ids the device holds per slot"""
        from .entities import RandomWaypoint, User
        ids = self._batch.ue_ids()[0]
        if [int(ue.id) for ue in self.ue_list] == ids.tolist():
            return
        ues = []
        for uid in ids.tolist():
            if uid not in self._ue_by_id:          # add_new_ue: 'slow' RandomWaypoint UE appearing on the map border
                self._ue_by_id[uid] = User(str(uid), self.map, None, None, movement=RandomWaypoint(self.map, velocity='slow'))
            ues.append(self._ue_by_id[uid])
        self.ue_list = ues

    def _reset_ue_list(self):
        self.ue_list = list(self.original_ue_list)                   # base.py:176-182

    def _step_batch(self, per_ue_actions):
        obs, reward, _, info = self._batch.step(self._device_actions(per_ue_actions), info=True)
        self._batch.check_errors()     # device-side flags (action range, waypoint table) raise here, not silently diverge
        if self._batch._dynamic:
            self._sync_ue_list()
        self.time += 1
        sum_utility = float(info['sum_utility'][0])
        self.total_utility += sum_utility
        self.last_lost_conn = info['lost_conn'][0].cpu().numpy()
        return (obs[0].cpu().numpy(), reward[0].cpu().numpy(), info['curr_dr'][0].cpu().numpy(),
                info['utility'][0].cpu().numpy(), sum_utility)


class CentralRelNormEnv(_MobileEnvFacade):
    """Single central agent controlling all UEs (reference central.py:9-73,143-152)."""
    _kind = 'central'

    def __init__(self, env_config):
        super().__init__(env_config)
        n, m = self.max_ues, self.num_bs
        self.action_space = spaces.MultiDiscrete([m + 1 for _ in range(n)])           # central.py:28
        self.observation_space = spaces.Dict({                                          # central.py:147-152
            'connected': spaces.MultiBinary(n * m),
            'dr': spaces.Box(low=0, high=1, shape=(n * m,)),
            'utility': spaces.Box(low=-1, high=1, shape=(n,)),
        })

    def _obs_dict(self, flat):
        """central.py:31-57: values are Python lists in the reference"""
        nm = self.max_ues * self.num_bs                     # zero-padded to max_ues (central.py:46-55)
        return {'connected': [int(v) for v in flat[:nm]], 'dr': [float(v) for v in flat[nm:2 * nm]],
                'utility': [float(v) for v in flat[2 * nm:]]}

    def get_ue_actions(self, action):
        """central.py:59-63"""
        assert self.action_space.contains(np.asarray(action)), \
            f"Action {action} does not fit action space {self.action_space}"
        return {ue: int(action[i]) for i, ue in enumerate(self.ue_list)}

    def reset(self):
        self.time = 0
        self._reset_ue_list()
        flat = self._batch.reset()[0].cpu().numpy()
        self.obs = self._obs_dict(flat)
        return self.obs

    def step(self, action):
        action = np.asarray(action)
        assert self.action_space.contains(action), f"Action {action} does not fit action space {self.action_space}"
        flat, reward, curr_dr, utility, sum_utility = self._step_batch(action.astype(np.int32)[:self.num_ue])
        self._sync_entities(curr_dr, utility)
        self.obs = self._obs_dict(flat)
        return self.obs, float(reward), self.done(), self._info(curr_dr, utility, sum_utility)


class CentralMaxNormEnv(CentralRelNormEnv):
    """CentralRelNormEnv with MaxNormEnv's SNR normalisation (reference central.py:155-164, variants.py:308-332; the
    alternative named in env_setup.py:35): obs['dr'] = (min(snr, 7e-6) - 2e-8) / (7e-6 - 2e-8), negative out of range."""
    _obs_norm = 'max'

    def __init__(self, env_config):
        super().__init__(env_config)
        n, m = self.max_ues, self.num_bs
        self.obs_space_dict['dr'] = spaces.Box(low=-1, high=1, shape=(m,))              # variants.py:313-317
        self.observation_space = spaces.Dict({                                          # central.py:159-164
            'connected': spaces.MultiBinary(n * m),
            'dr': spaces.Box(low=-1, high=1, shape=(n * m,)),
            'utility': spaces.Box(low=-1, high=1, shape=(n,)),
        })


class _CentralVariantEnv(CentralRelNormEnv):
    """central env whose observation is one of the data-rate classes: a dict of the keys present, each a Python list"""

    def _obs_dict(self, flat):
        out, o = {}, 0
        for key, w in self._batch.obs_keys:
            vals = flat[o:o + w]
            out[key] = [int(v) for v in vals] if key in ('connected', 'ues_at_bs') else [float(v) for v in vals]
            o += w
        return out


class CentralNormDrEnv(_CentralVariantEnv):
    """Reference central.py:107-140 over NormDrMobileEnv (variants.py:173-250): every UE observes the shared rate it gets
    or would get from every BS, cut at 100 and normalised; `dr_total` is its current total rate."""
    _obs_variant = 'normdr'

    def __init__(self, env_config):
        super().__init__(env_config)
        n, m = self.max_ues, self.num_bs
        self.dr_cutoff = 100
        self.observation_space = spaces.Dict({                                          # central.py:120-131
            'dr': spaces.Box(low=0, high=1, shape=(n * m,)),
            'connected': spaces.MultiBinary(n * m),
            'dr_total': spaces.Box(low=0, high=1, shape=(n,)),
        })


class CentralDrEnv(_CentralVariantEnv):
    """Reference central.py:75-104 over DatarateMobileEnv (variants.py:42-170); extra env_config keys: dr_cutoff ('auto' or
    a number), sub_req_dr, curr_dr_obs, ues_at_bs_obs, dist_obs, next_dist_obs."""
    _obs_variant = 'datarate'

    def _obs_options(self):
        ec = self.env_config
        return {k: ec[k] for k in ('dr_cutoff', 'sub_req_dr', 'curr_dr_obs', 'ues_at_bs_obs', 'dist_obs', 'next_dist_obs')}

    def __init__(self, env_config):
        super().__init__(env_config)
        n, m = self.max_ues, self.num_bs
        o = self._batch.obs_opts
        self.dr_cutoff, self.sub_req_dr = o['dr_cutoff'], o['sub_req_dr']
        self.curr_dr_obs, self.ues_at_bs_obs = o['curr_dr_obs'], o['ues_at_bs_obs']
        self.dist_obs, self.next_dist_obs = o['dist_obs'], o['next_dist_obs']
        obs_space = {'dr': spaces.Box(low=-1, high=1, shape=(n * m,)), 'connected': spaces.MultiBinary(n * m)}   # central.py:88-92
        if self.curr_dr_obs:
            obs_space['dr_total'] = spaces.Box(low=-1, high=1, shape=(n,))
        if self.ues_at_bs_obs:
            obs_space['ues_at_bs'] = spaces.MultiDiscrete([n + 1 for _ in range(m)])
        if self.dist_obs:
            obs_space['dist'] = spaces.Box(low=0, high=1, shape=(n * m,))
        if self.next_dist_obs:
            obs_space['next_dist'] = spaces.Box(low=0, high=1, shape=(n * m,))
        self.observation_space = spaces.Dict(obs_space)


class MultiAgentMobileEnv(_MobileEnvFacade):
    """One agent per UE, RLlib MultiAgentEnv protocol (reference multi_agent.py:6-107)."""
    _kind = 'multi'

    def __init__(self, env_config):
        super().__init__(env_config)
        self.action_space = self._ue_action_space                                       # variants.py:17
        self.observation_space = spaces.Dict(self.obs_space_dict)                       # variants.py:269

    def _obs_dict(self, packed):
        m = self.num_bs
        out = {}
        for i, ue in enumerate(self.ue_list):
            row = packed[i]
            out[ue.id] = {'connected': [int(v) for v in row[:m]], 'dr': [float(v) for v in row[m:2 * m]],
                          'utility': [float(row[4 * m])], 'ues_at_bs': [float(v) for v in row[2 * m:3 * m]],
                          'util_at_bs': [float(v) for v in row[3 * m:4 * m]]}
        return out

    def get_ue_actions(self, action):
        """multi_agent.py:21-30: UEs missing from the dict do nothing"""
        return {ue: action[ue.id] for ue in self.ue_list if ue.id in action}

    def reset(self):
        self.time = 0
        self._reset_ue_list()
        packed = self._batch.reset()[0].cpu().numpy()
        self.obs = self._obs_dict(packed)
        return self.obs

    def done(self):
        """multi_agent.py:97-102"""
        dones = {ue.id: None for ue in self.ue_list}
        dones['__all__'] = None
        return dones

    def step(self, action):
        per_ue = np.zeros(self.num_ue, dtype=np.int32)
        for i, ue in enumerate(self.ue_list):
            if ue.id in action:
                a = action[ue.id]
                assert self.action_space.contains(int(a)), f"Action {a} does not fit action space {self.action_space}"
                per_ue[i] = int(a)
        packed, reward, curr_dr, utility, sum_utility = self._step_batch(per_ue)
        self._sync_entities(curr_dr, utility)
        self.obs = self._obs_dict(packed)
        rewards = {ue.id: float(reward[i]) for i, ue in enumerate(self.ue_list)}         # multi_agent.py:39-95
        info = self._info(curr_dr, utility, sum_utility)
        return self.obs, rewards, self.done(), {ue.id: info for ue in self.ue_list}      # multi_agent.py:104-107


class SeqMultiAgentMobileEnv(MultiAgentMobileEnv):
    """
    All agents observe and act one after the other within a time step; the UEs move and time advances after the last one
    (reference multi_agent.py:110-179).  Every return value holds the CURRENT UE only.  As in the reference, `done` and
    `info` nest the parent's per-UE dicts under the current UE's id (multi_agent.py:130-143 call the parent's methods,
    which already return dicts keyed by UE id), and `ue_order_idx` survives reset().
    """

    def __init__(self, env_config):
        super().__init__(env_config)
        self.ue_order = self.ue_list
        self.ue_order_idx = 0
        self.curr_ue = self.ue_order[self.ue_order_idx]
        self._last_info = None

    def _curr_obs(self, row):
        m = self.num_bs
        return {self.curr_ue.id: {'connected': [int(v) for v in row[:m]], 'dr': [float(v) for v in row[m:2 * m]],
                                  'utility': [float(row[4 * m])], 'ues_at_bs': [float(v) for v in row[2 * m:3 * m]],
                                  'util_at_bs': [float(v) for v in row[3 * m:4 * m]]}}

    def reset(self):
        self.time = 0
        self._reset_ue_list()
        packed = self._batch.reset()[0].cpu().numpy()
        self.curr_ue = self.ue_order[self.ue_order_idx]
        self.obs = self._curr_obs(packed[self.ue_order_idx])
        return self.obs

    def done(self):
        done = super().done()
        return {self.curr_ue.id: done, '__all__': done}

    def step(self, action):
        a = 0
        if self.curr_ue.id in action:                                   # multi_agent.py:21-30
            a = int(action[self.curr_ue.id])
            assert self.action_space.contains(a), f"Action {a} does not fit action space {self.action_space}"
        dev = self._batch.device
        row, reward, _, info = self._batch.step_sequential(torch.as_tensor([a], dtype=torch.int32, device=dev), info=True)
        self._batch.check_errors()     # device-side flags raise here, as in _step_batch
        if info['moved']:
            self.time += 1
        self.ue_order_idx = info['ue_index']
        self.curr_ue = self.ue_order[self.ue_order_idx]
        curr_dr, utility = info['curr_dr'][0].cpu().numpy(), info['utility'][0].cpu().numpy()
        sum_utility = float(info['sum_utility'][0])
        self.last_lost_conn = info['lost_conn'][0].cpu().numpy()
        self._sync_entities(curr_dr, utility)
        self.obs = self._curr_obs(row[0].cpu().numpy())
        base_info = self._info(curr_dr, utility, sum_utility)
        return (self.obs, {self.curr_ue.id: float(reward[0])}, self.done(),
                {self.curr_ue.id: {ue.id: base_info for ue in self.ue_list}})


def get_env_class(env_type):
    """deepcomp/util/env_setup.py:23-37"""
    if env_type == 'central':
        return CentralRelNormEnv
    if env_type == 'multi':
        return MultiAgentMobileEnv
    raise NotImplementedError(f"env type {env_type!r}: only 'central' and 'multi' are on the B200 hot path")
