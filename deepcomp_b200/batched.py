"""Batched tensor surface of the B200 env step: K independent DeepCoMP env instances advanced by one CUDA launch.

This is the layer the gym / RLlib-shaped facades (``deepcomp_b200.env``, ``deepcomp_b200.rllib``) sit on.  It mirrors
``MobileEnv`` (reference deepcomp/env/single_ue/base.py:20-466) for K envs at once: ``reset() -> obs``,
``step(actions) -> (obs, reward, done, info)``, ``seed``-able per env, with ``done`` always ``None``
(base.py:371-381).  PyTorch only provides device memory and the stream; all compute is in libdeepcomp_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import DcbConfig, DcbObsVariant, DcbOutputs, DcbPolicy, DcbStateHost, check

KIND = {'central': 0, 'multi': 1}
REWARD = {'avg': 0, 'sum': 1, 'min': 2}
SHARING = {'resource-fair': 0, 'rate-fair': 1, 'proportional-fair': 2, 'max-cap': 3}
SHARING_MIX = ['resource-fair', 'rate-fair', 'proportional-fair']   # reference util/env_setup.py:48
VELOCITY = {'slow': -1.0, 'fast': -2.0}


def sharing_for_bs(sharing, bs_idx):
    """reference util/env_setup.py:40-49 get_sharing_for_bs"""
    if sharing != 'mixed':
        if sharing not in SHARING:
            raise ValueError(f"sharing model {sharing!r} not supported; one of {sorted(SHARING)} or 'mixed'")
        return sharing
    return SHARING_MIX[bs_idx % len(SHARING_MIX)]


def env_seeds(base_seed, num_envs, n_ue, first_env=0):
    """
    Seeds for a batch of envs.  UE i of an env draws from random.Random(seed + 100*i) (base.py:138-143), so consecutive
    env seeds must be >= 100*(N+1) apart or UE streams of different envs collide.
    `first_env` is the global index of this shard's first env, so a sharded batch reproduces the unsharded one.
    """
    k = np.arange(first_env, first_env + num_envs, dtype=np.int64)
    return np.int64(base_seed) + k * np.int64(100 * (n_ue + 1))


class BatchedMobileEnv:
    """K env instances of one scenario (same map, BS layout and UE classes; per-env seeds) on one GPU."""

    def __init__(self, num_envs, n_ue, bs_xy, map_wh, kind='multi', sharing='mixed', velocities='slow', seed=0,
                 seeds=None, reward='avg', episode_length=100, rand_episodes=False, auto_reset=False, init_pos=None,
                 pause_duration=2, border_buffer=10, device=None, first_env=0, max_ues=None, ue_arrival=None,
                 new_ue_interval=None, util_func='log', dr_req=1, obs_norm='rel', uniform_moves=None,
                 obs_variant=None, obs_opts=None, interference=False):
        """
        `obs_norm`: 'rel' = the observation entry 'dr' is snr / max snr (RelNormEnv, variants.py:276-284; default);
        'max' = (min(snr, 7e-6) - 2e-8) / (7e-6 - 2e-8) (MaxNormEnv, variants.py:308-332; CentralMaxNormEnv).

        `obs_variant`: None, 'normdr' (CentralNormDrEnv, central.py:107-140) or 'datarate' (CentralDrEnv, central.py:75-104)
        with `obs_opts` = the reference's env_config keys dr_cutoff ('auto' or a number), sub_req_dr, curr_dr_obs,
        ues_at_bs_obs, dist_obs, next_dist_obs (variants.py:56-79); central kind only.
        `interference`: extension, not in the reference (SNR only): SINR = P_b / (noise + sum of the other P_b').

        `uniform_moves`: None, or per UE None (RandomWaypoint) / (move_x, move_y) = UniformMovement (util/movement.py:26-80),
        each component a number or 'slow' / 'fast' (drawn per reset from the UE's movement generator).

        Variable UE population (reference env_config keys of the same names, base.py:80-84, 429-443): `n_ue` UEs at
        reset, `max_ues` slots per env (every per-UE array has max_ues rows; rows of UEs that are not there read as
        zeros, central.py:46-55), `ue_arrival` = {time: +n arrivals / -n departures} (env_setup.py:205-226),
        `new_ue_interval` = one arrival every so many steps.  All envs of the batch step in lockstep.
        """
        if not torch.cuda.is_available():
            raise RuntimeError("deepcomp_b200 needs a CUDA device: the env step is a CUDA kernel and has no CPU "
                               "fallback")
        self._L = _lib.load()
        if kind not in KIND:
            raise ValueError(f"kind must be 'central' or 'multi', got {kind!r}")
        if reward not in REWARD:
            # reference raises NotImplementedError for unknown aggregations (central.py:73, multi_agent.py:92)
            raise NotImplementedError(f"Unexpected reward aggregation: {reward}")
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.num_ue_initial = int(n_ue)
        if max_ues is not None and int(max_ues) < int(n_ue):
            raise ValueError(f"max_ues = {max_ues} < n_ue = {n_ue}")             # base.py:84
        n_slots = int(n_ue) if max_ues is None else int(max_ues)
        self.ue_arrival = None if ue_arrival is None else {int(t): int(n) for t, n in dict(ue_arrival).items()}
        self.new_ue_interval = None if new_ue_interval is None else int(new_ue_interval)
        self._dynamic = self.ue_arrival is not None or self.new_ue_interval is not None
        if self._dynamic and (rand_episodes or auto_reset):
            raise NotImplementedError("variable UE population needs rand_episodes=False and auto_reset=False")
        self._t = 0                                          # MobileEnv.time of the lockstep batch (events key on it)
        self._table_steps = 0                                # steps taken since the waypoint tables were (re)generated
        if not isinstance(velocities, (list, tuple, np.ndarray)):
            velocities = [velocities] * int(n_ue)
        velocities = list(velocities) + ['slow'] * (n_slots - len(velocities))   # add_new_ue(velocity='slow')
        if init_pos is not None:
            init_pos = list(init_pos) + [('random', 'random')] * (n_slots - len(init_pos))
        n_ue = n_slots                                       # from here on n_ue counts SLOTS (= max_ues)
        self.num_envs, self.n_ue, self.kind, self.reward_agg = int(num_envs), int(n_ue), kind, reward
        bs = np.ascontiguousarray(np.asarray(bs_xy, dtype=np.float64).reshape(-1, 2))
        self.n_bs = bs.shape[0]
        self.bs_xy = bs
        self.map_wh = (int(map_wh[0]), int(map_wh[1]))      # Map casts to int (map.py:20-21)
        self.episode_length = int(episode_length)
        self.rand_episodes, self.auto_reset = bool(rand_episodes), bool(auto_reset)
        if isinstance(sharing, str):
            sharing = [sharing_for_bs(sharing, b) for b in range(self.n_bs)]
        for s in sharing:
            if s not in SHARING:
                raise ValueError(f"sharing model {s!r} not supported; one of {sorted(SHARING)}")
        self.sharing = list(sharing)
        if not isinstance(velocities, (list, tuple, np.ndarray)):
            velocities = [velocities] * self.n_ue
        if len(velocities) != self.n_ue:
            raise ValueError("need one velocity per UE")
        self.velocities = list(velocities)
        vel = np.array([VELOCITY[v] if isinstance(v, str) else float(v) for v in velocities], dtype=np.float64)
        ixy = np.full((self.n_ue, 2), np.nan, dtype=np.float64)
        if init_pos is not None:
            for i, (px, py) in enumerate(init_pos):
                if px != 'random':
                    ixy[i, 0] = float(px)
                if py != 'random':
                    ixy[i, 1] = float(py)
        if seeds is None:
            # arriving UEs get ids (and seeds 100 * id apart) beyond the slots: keep the env seeds further apart
            n_ids = self.n_ue + (sum(max(n, 0) for n in self.ue_arrival.values()) if self.ue_arrival else 0) + \
                (self.episode_length // self.new_ue_interval if self.new_ue_interval else 0)
            seeds = env_seeds(seed, self.num_envs, n_ids, first_env)
        seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.int64))
        if seeds.shape != (self.num_envs,):
            raise ValueError("need one seed per env")
        self.seeds = seeds
        sh = np.array([SHARING[s] for s in self.sharing], dtype=np.int32)

        cfg = DcbConfig(
            abi_version=_lib.DCB_ABI_VERSION, device=self.device.index, kind=KIND[kind], reward=REWARD[reward],
            num_envs=self.num_envs, n_ue=self.n_ue, n_bs=self.n_bs, map_width=self.map_wh[0],
            map_height=self.map_wh[1], episode_length=self.episode_length, rand_episodes=int(self.rand_episodes),
            auto_reset=int(self.auto_reset), pause_duration=int(pause_duration), border_buffer=int(border_buffer),
            host_bs_xy=bs.ctypes.data, host_sharing=sh.ctypes.data, host_velocity=vel.ctypes.data,
            host_init_xy=ixy.ctypes.data, host_seeds=seeds.ctypes.data)
        h = ctypes.c_void_p()
        check(self._L.dcb_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h
        self.obs_size = int(self._L.dcb_obs_size(h))
        self.reward_size = int(self._L.dcb_reward_size(h))
        self.obs_shape = (self.obs_size,) if kind == 'central' else (self.n_ue, 4 * self.n_bs + 1)
        self.reward_shape = () if kind == 'central' else (self.n_ue,)
        self._pinned = None
        if self.num_ue_initial != self.n_ue:
            self.active_ues = self.num_ue_initial
        if util_func not in ('log', 'step'):
            # the reference's 'linear' utility asserts MIN_UTILITY == 0 (utility.py:18): unusable with its own constants
            raise NotImplementedError(f"Utility function {util_func} not implemented!")           # user.py:92
        self.util_func = util_func
        if util_func != 'log':
            check(self._L.dcb_set_utility(self._h, 1, float(dr_req)))
        if obs_norm not in ('rel', 'max'):
            raise ValueError(f"obs_norm must be 'rel' (RelNormEnv) or 'max' (MaxNormEnv), got {obs_norm!r}")
        self.obs_norm = obs_norm
        if obs_norm == 'max':
            check(self._L.dcb_set_obs_norm(self._h, 1))
        self.uniform_moves = None
        if uniform_moves is not None and any(u is not None for u in uniform_moves):
            if self._dynamic or self.num_ue_initial != self.n_ue:
                raise NotImplementedError("UniformMovement UEs with a variable UE population")
            if len(uniform_moves) != self.n_ue:
                raise ValueError("need one uniform_moves entry (None or (move_x, move_y)) per UE")
            kinds = np.zeros((self.n_ue, 2), dtype=np.int32)
            vals = np.zeros((self.n_ue, 2), dtype=np.float64)
            for i, u in enumerate(uniform_moves):
                if u is None:
                    continue
                for c, m in enumerate(u):
                    if m in ('slow', 'fast'):
                        kinds[i, c] = 2 if m == 'slow' else 3           # movement.py:49-52: randint(1, 5) / randint(10, 20)
                    else:
                        kinds[i, c], vals[i, c] = 1, float(m)
            check(self._L.dcb_set_uniform_movement(self._h, ctypes.c_void_p(kinds.ctypes.data),
                                                   ctypes.c_void_p(vals.ctypes.data)))
            self.uniform_moves = [None if u is None else tuple(u) for u in uniform_moves]
        self.obs_variant, self.obs_opts = obs_variant, None
        if obs_variant is not None:
            if obs_variant not in ('normdr', 'datarate'):
                raise ValueError(f"obs_variant must be None, 'normdr' or 'datarate', got {obs_variant!r}")
            if kind != 'central':
                raise NotImplementedError("the data-rate observation classes exist for the central env only "
                                          "(CentralNormDrEnv / CentralDrEnv, central.py:75-140)")
            o = dict(dr_cutoff='auto', sub_req_dr=True, curr_dr_obs=False, ues_at_bs_obs=False, dist_obs=False,
                     next_dist_obs=False)
            o.update(obs_opts or {})
            v = DcbObsVariant(kind=1 if obs_variant == 'normdr' else 2)
            if obs_variant == 'datarate':
                # variants.py:75-79
                assert not (o['dr_cutoff'] == 'auto' and not o['sub_req_dr']), "For dr_cutoff auto, sub_req_dr must be True."
                assert (not o['curr_dr_obs']) or (o['dr_cutoff'] == 'auto' and o['sub_req_dr']), \
                    "Enable all processing to add extra obs"
                assert o['dist_obs'] or not o['next_dist_obs'], "Also enable 'dist_obs' when using 'next_dist_obs'"
                v.dr_mode = 0 if o['dr_cutoff'] == 'auto' else (1 if o['sub_req_dr'] else 2)
                v.dr_cutoff = 0.0 if o['dr_cutoff'] == 'auto' else float(o['dr_cutoff'])
                v.curr_dr_obs, v.ues_at_bs_obs = int(bool(o['curr_dr_obs'])), int(bool(o['ues_at_bs_obs']))
                v.dist_obs, v.next_dist_obs = int(bool(o['dist_obs'])), int(bool(o['next_dist_obs']))
                if util_func == 'log':          # the required rate enters the observation even with the log utility
                    check(self._L.dcb_set_utility(self._h, 0, float(dr_req)))
            check(self._L.dcb_set_obs_variant(self._h, ctypes.byref(v)))
            self.obs_opts = o
            self.obs_size = int(self._L.dcb_obs_size(self._h))
            self.obs_shape = (self.obs_size,)
            n, m = self.n_ue, self.n_bs
            if obs_variant == 'normdr':
                self.obs_keys = [('connected', n * m), ('dr', n * m), ('dr_total', n)]
            else:
                self.obs_keys = [('connected', n * m)] + ([('dist', n * m)] if o['dist_obs'] else []) + [('dr', n * m)] + \
                    ([('dr_total', n)] if o['curr_dr_obs'] else []) + ([('next_dist', n * m)] if o['next_dist_obs'] else []) + \
                    ([('ues_at_bs', n * m)] if o['ues_at_bs_obs'] else [])
            assert sum(w for _, w in self.obs_keys) == self.obs_size
        self.interference = bool(interference)
        if self.interference:
            check(self._L.dcb_set_interference(self._h, 1))
        self._seq_idx = 0            # SeqMultiAgentMobileEnv.ue_order_idx (multi_agent.py:119; never reset)

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self._L.dcb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    @property
    def algorithmic_bytes_per_env_step(self):
        return int(self._L.dcb_algorithmic_bytes_per_env_step(self._h))

    @property
    def launch_count(self):
        return int(self._L.dcb_launch_count(self._h))

    @property
    def active_ues(self):
        """UEs present per env: slots [0, active_ues) of the n_ue = max_ues slots (reference base.py:80-84)"""
        return int(self._L.dcb_get_active_ues(self._h))

    @active_ues.setter
    def active_ues(self, n):
        check(self._L.dcb_set_active_ues(self._h, int(n)))

    # ------------------------------------------------------------------ brute force (reference agent/brute_force.py)
    @property
    def num_joint_actions(self):
        """(M + 1)^num_ue candidates of BruteForceAgent (brute_force.py:86-88)"""
        return int(self._L.dcb_num_joint_actions(self._h))

    def candidate_action(self, c):
        """brute_force.py:26-62: digits of c in base M + 1, most significant digit = first UE; zeros on padding slots"""
        digits = []
        for _ in range(self.active_ues):
            digits.append(c % (self.n_bs + 1))
            c //= self.n_bs + 1
        return digits[::-1] + [0] * (self.n_ue - self.active_ues)

    def test_actions(self, env_index=0, first=0, count=None):
        """
        MobileEnv.test_ue_actions (base.py:284-313) + central step_reward (central.py:65-73) for the joint actions
        [first, first + count) of one env, all at once on the device; the env state is not changed.  float64 [count].
        """
        if count is None:
            count = self.num_joint_actions - first
        out = self._empty((int(count),), torch.float64)
        check(self._L.dcb_test_actions(self._h, int(env_index), int(first), int(count), ctypes.c_void_p(out.data_ptr()),
                                       self._stream()))
        return out

    def best_joint_action(self, env_index=0, chunk=1 << 24):
        """BruteForceAgent.compute_action (brute_force.py:79-94): the first maximum over all joint actions"""
        best, best_c = -float('inf'), 0
        for first in range(0, self.num_joint_actions, chunk):
            r = self.test_actions(env_index, first, min(chunk, self.num_joint_actions - first))
            mx = float(r.max())
            if mx > best:
                best, best_c = mx, first + int((r == mx).nonzero()[0, 0])
        return self.candidate_action(best_c), best

    def ue_ids(self):
        """User.id (as integers) of the UEs present, int32 [K, active_ues]; ids of arrivals continue after the last id"""
        ids = np.zeros((self.num_envs, self.n_ue), dtype=np.int32)
        check(self._L.dcb_get_ue_ids(self._h, ctypes.c_void_p(ids.ctypes.data)))
        return ids[:, :self.active_ues]

    @property
    def kernel_name(self):
        """'dcb_step_kernel' (fused, several envs per CTA) or 'dcb_wide_kernel' (one CTA per env, large envs)"""
        return self._L.dcb_kernel_name(self._h).decode()

    @property
    def launch_geometry(self):
        v = [ctypes.c_int32() for _ in range(4)]
        check(self._L.dcb_launch_geometry(self._h, *[ctypes.byref(x) for x in v]))
        return dict(envs_per_cta=v[0].value, threads=v[1].value, smem_bytes=v[2].value, grid=v[3].value)

    def check_errors(self):
        """Raise ValueError if an out-of-range action reached the device (reference: assert in base.py:238 /
        central.py:61), RuntimeError if a UE ran out of pre-drawn waypoints."""
        rc = self._L.dcb_check_errors(self._h, self._stream())
        if rc == -4:
            raise ValueError(self._L.dcb_last_error().decode())
        if rc != 0:
            raise RuntimeError(self._L.dcb_last_error().decode())

    # ------------------------------------------------------------------ observation helpers
    def split_obs(self, obs):
        """Packed observation -> dict of views keyed like the reference's obs dicts (variants.py:302-303)."""
        M, N = self.n_bs, self.n_ue
        if self.obs_variant is not None:
            out, o = {}, 0
            for key, w in self.obs_keys:
                out[key] = obs[..., o:o + w]
                o += w
            return out
        if self.kind == 'central':
            nm = N * M
            return {'connected': obs[..., :nm], 'dr': obs[..., nm:2 * nm], 'utility': obs[..., 2 * nm:]}
        return {'connected': obs[..., :M], 'dr': obs[..., M:2 * M], 'ues_at_bs': obs[..., 2 * M:3 * M],
                'util_at_bs': obs[..., 3 * M:4 * M], 'utility': obs[..., 4 * M:]}

    def _outputs(self, T, obs=True, info=False, debug=False):
        K, N, M = self.num_envs, self.n_ue, self.n_bs
        lead = (T,) if T else ()
        o = DcbOutputs()
        t = {}
        if obs:
            t['obs'] = self._empty(lead + (K,) + self.obs_shape, torch.float32)
            o.obs, o.obs_stride = t['obs'].data_ptr(), K * self.obs_size
        if T != 0:
            t['reward'] = self._empty(lead + (K,) + self.reward_shape, torch.float32)
            o.reward, o.reward_stride = t['reward'].data_ptr(), K * self.reward_size
            t['lost_conn'] = self._empty(lead + (K, N), torch.uint8)
            o.lost_conn, o.lost_conn_stride = t['lost_conn'].data_ptr(), K * N
        if info:
            t['curr_dr'] = self._empty(lead + (K, N), torch.float32)
            t['utility'] = self._empty(lead + (K, N), torch.float32)
            t['sum_utility'] = self._empty(lead + (K,), torch.float32)
            o.curr_dr, o.curr_dr_stride = t['curr_dr'].data_ptr(), K * N
            o.utility, o.utility_stride = t['utility'].data_ptr(), K * N
            o.sum_utility, o.sum_utility_stride = t['sum_utility'].data_ptr(), K
        if debug:
            f64 = torch.float64
            t['dbg_obs'] = self._empty((K,) + self.obs_shape, f64)
            t['dbg_snr'] = self._empty((K, N, M), f64)
            t['dbg_link_rate'] = self._empty((K, N, M), f64)
            t['dbg_curr_dr'] = self._empty((K, N), f64)
            t['dbg_utility'] = self._empty((K, N), f64)
            t['dbg_sum_utility'] = self._empty((K,), f64)
            o.dbg_obs, o.dbg_snr = t['dbg_obs'].data_ptr(), t['dbg_snr'].data_ptr()
            o.dbg_link_rate, o.dbg_curr_dr = t['dbg_link_rate'].data_ptr(), t['dbg_curr_dr'].data_ptr()
            o.dbg_utility, o.dbg_sum_utility = t['dbg_utility'].data_ptr(), t['dbg_sum_utility'].data_ptr()
            if T != 0:
                t['dbg_reward'] = self._empty((K,) + self.reward_shape, f64)
                o.dbg_reward = t['dbg_reward'].data_ptr()
        return o, t

    # ------------------------------------------------------------------ gym-like batched API
    def _population_events(self, t, actions):
        """base.py:429-443 at MobileEnv.time == t: arrivals / departures between the actions and the rate update.
        Returns the action tensor to step with (a copy, edited on the device, if something happened)."""
        ev = []
        if self.new_ue_interval is not None and t > 0 and t % self.new_ue_interval == 0:
            ev.append((1, 0))
        if self.ue_arrival is not None and t in self.ue_arrival:
            n = self.ue_arrival[t]
            ev.append((n, 0) if n > 0 else (0, -n))
        if ev:
            actions = actions.clone()
            for n_add, n_rem in ev:
                check(self._L.dcb_population_event(self._h, n_add, n_rem, ctypes.c_void_p(actions.data_ptr()),
                                                   self._stream()))
        return actions

    def _before_steps(self, T):
        """
        Continuous stepping (the reference's --cont-train / soft_horizon: no reset at episode_length, `done` is never set,
        base.py:371-381): the pre-drawn waypoint tables cover episode_length steps, so before a launch would run past them
        the per-UE random streams are continued on the device (dcb_extend_waypoints).
        """
        if self.auto_reset:                # the envs reset themselves on the device when their time is up
            return
        if self._table_steps + T <= self.episode_length:
            self._table_steps += T
            return
        if T > self.episode_length:
            raise ValueError(f"a fragment of {T} steps is longer than episode_length = {self.episode_length}")
        if self._dynamic or self.num_ue_initial != self.n_ue:
            raise NotImplementedError("stepping past episode_length without reset() with a variable UE population")
        check(self._L.dcb_extend_waypoints(self._h, self._stream()))
        self._table_steps = T

    def reset(self, env_ids=None, debug=False):
        """MobileEnv.reset (base.py:169-189) for all envs or the listed ones; returns the observation of ALL envs."""
        if self._dynamic and env_ids is not None:
            raise NotImplementedError("a batch with a variable UE population resets as a whole (lockstep)")
        self._t = 0
        if env_ids is None:
            self._table_steps = 0
        if env_ids is None:
            check(self._L.dcb_reset(self._h, None, 0, self._stream()))
        else:
            ids = np.ascontiguousarray(np.asarray(env_ids, dtype=np.int32))
            check(self._L.dcb_reset(self._h, ctypes.c_void_p(ids.ctypes.data), ids.size, self._stream()))
        return self.observe(debug=debug)

    def observe(self, debug=False):
        o, t = self._outputs(0, debug=debug, info=debug)
        check(self._L.dcb_observe(self._h, ctypes.byref(o), self._stream()))
        return t if debug else t['obs']

    def _check_actions(self, actions, T=None):
        shape = (self.num_envs, self.n_ue) if T is None else (T, self.num_envs, self.n_ue)
        if not (isinstance(actions, torch.Tensor) and actions.is_cuda and actions.dtype == torch.int32
                and actions.is_contiguous() and tuple(actions.shape) == shape and actions.device == self.device):
            raise ValueError(f"actions must be a contiguous int32 tensor of shape {shape} on {self.device}")

    def step(self, actions, info=True, debug=False):
        """MobileEnv.step (base.py:413-466) for all K envs.  actions: int32 [K, N] on the device."""
        self._check_actions(actions)
        self._before_steps(1)
        if self._dynamic:
            actions = self._population_events(self._t, actions)
        self._t += 1
        o, t = self._outputs(None, info=info or debug, debug=debug)
        check(self._L.dcb_step(self._h, ctypes.c_void_p(actions.data_ptr()), ctypes.byref(o), self._stream()))
        if debug:
            return t
        inf = {'lost_conn': t['lost_conn']}
        if info:
            inf.update(curr_dr=t['curr_dr'], utility=t['utility'], sum_utility=t['sum_utility'])
        return t['obs'], t['reward'], None, inf

    def step_sequential(self, action, info=True, debug=False):
        """
        SeqMultiAgentMobileEnv.step (multi_ue/multi_agent.py:149-179) for all K envs: only the CURRENT UE of the round
        acts (`action`: int32 [K] on the device, or [K, N] of which the current UE's column is used); rates and rewards
        are updated; after the last UE of the round the UEs move and time advances (a plain step), otherwise nothing
        moves (dcb_step_no_move).  Returns the observation row and the reward of the NEXT UE: (obs [K, 4M+1],
        reward [K], None, info) with info['ue_index'] = that UE's slot.
        """
        if self.kind != 'multi' or self._dynamic:
            raise NotImplementedError("sequential stepping is SeqMultiAgentMobileEnv: multi-agent, fixed UE population")
        K, N = self.num_envs, self.n_ue
        cur = self._seq_idx
        a = torch.zeros((K, N), dtype=torch.int32, device=self.device)
        a[:, cur] = action if action.dim() == 1 else action[:, cur]
        o, t = self._outputs(None, info=info or debug, debug=debug)
        last_of_round = cur + 1 >= self.active_ues
        if last_of_round:
            self._before_steps(1)
            self._t += 1
            check(self._L.dcb_step(self._h, ctypes.c_void_p(a.data_ptr()), ctypes.byref(o), self._stream()))
        else:
            check(self._L.dcb_step_no_move(self._h, ctypes.c_void_p(a.data_ptr()), ctypes.byref(o), self._stream()))
        self._seq_idx = 0 if last_of_round else cur + 1
        nxt = self._seq_idx
        if debug:
            t['ue_index'] = nxt
            return t
        inf = {'lost_conn': t['lost_conn'], 'ue_index': nxt, 'moved': last_of_round}
        if info:
            inf.update(curr_dr=t['curr_dr'], utility=t['utility'], sum_utility=t['sum_utility'])
        return t['obs'][:, nxt], t['reward'][:, nxt], None, inf

    def step_many(self, actions, obs=True, info=False, out=None):
        """
        T consecutive steps in one launch.  actions: int32 [T, K, N].  Returns dict of [T, ...] tensors
        (obs, reward, lost_conn[, curr_dr, utility, sum_utility]); pass `out` (from a previous call) to reuse buffers.
        """
        T = int(actions.shape[0])
        self._check_actions(actions, T)
        if out is None:
            o, t = self._outputs(T, obs=obs, info=info)
            t['_struct'] = o
        else:
            o, t = out['_struct'], out
        if self._dynamic:
            # arrivals / departures cut the fragment: one launch per stretch of steps without an event
            t0 = 0
            while t0 < T:
                a0 = self._population_events(self._t, actions[t0])
                t1 = t0 + 1
                while t1 < T and not self._has_event(self._t + (t1 - t0)):
                    t1 += 1
                acts = actions[t0:t1] if a0.data_ptr() == actions[t0].data_ptr() else \
                    torch.cat([a0[None], actions[t0 + 1:t1]]).contiguous()
                self._before_steps(t1 - t0)
                check(self._L.dcb_step_many(self._h, ctypes.c_void_p(acts.data_ptr()), t1 - t0,
                                            ctypes.byref(self._offset_outputs(o, t0)), self._stream()))
                self._t += t1 - t0
                t0 = t1
            return t
        self._t += T
        L = T if self.auto_reset else self.episode_length
        for t0 in range(0, T, L):                       # fragments longer than an episode: one launch per <= L steps
            n = min(L, T - t0)
            self._before_steps(n)
            check(self._L.dcb_step_many(self._h, ctypes.c_void_p(actions[t0:t0 + n].data_ptr()), n,
                                        ctypes.byref(o if t0 == 0 else self._offset_outputs(o, t0)), self._stream()))
        return t

    def _has_event(self, t):
        return (self.new_ue_interval is not None and t > 0 and t % self.new_ue_interval == 0) or \
            (self.ue_arrival is not None and self.ue_arrival.get(t, 0) != 0)

    @staticmethod
    def _offset_outputs(o, t0):
        """the outputs struct of a [T, ...] buffer set, advanced to step t0"""
        q = DcbOutputs()
        ctypes.memmove(ctypes.byref(q), ctypes.byref(o), ctypes.sizeof(DcbOutputs))
        for name, size in (('obs', 4), ('reward', 4), ('lost_conn', 1), ('curr_dr', 4), ('utility', 4), ('sum_utility', 4)):
            ptr = getattr(q, name)
            if ptr:
                setattr(q, name, ptr + t0 * getattr(q, name + '_stride') * size)
        return q

    def rollout(self, policy, T, obs=True, info=False, out=None, return_actions=True):
        """
        T consecutive steps driven by a scripted baseline policy ON THE DEVICE (no host round trip per step).
        `policy`: an agent from deepcomp_b200.agents (anything with .device_policy()) or the dict it returns.
        Returns the same dict as step_many plus 'actions' int32 [T, K, N] (the actions the policy took).
        """
        from .agents import POLICY_KIND
        if self._dynamic:
            raise NotImplementedError("device-side policies with a variable UE population")
        self._before_steps(int(T))
        self._t += T
        spec = policy.device_policy(self) if hasattr(policy, 'device_policy') else dict(policy)
        pol = DcbPolicy(kind=POLICY_KIND[spec['kind']], noop_interval=int(spec.get('noop_interval', 0)),
                        epsilon=float(spec.get('epsilon', 0.0)), seed=int(spec.get('seed', 0)),
                        calls_before=int(spec.get('calls_before', -1)))
        if hasattr(policy, 'device_calls'):          # the agent keeps its own call count (not the handle)
            policy.device_calls += int(T)
        keep = []
        if spec['kind'] == 'static':
            cm = np.ascontiguousarray(np.asarray(spec['cluster_masks'], dtype=np.uint64))
            if cm.shape != (self.n_bs,):
                raise ValueError("need one cluster mask per BS")
            keep.append(cm)
            pol.host_cluster_masks = cm.ctypes.data
        if spec['kind'] == 'fixed':
            fa = np.ascontiguousarray(np.asarray(spec['fixed_action'], dtype=np.int32))
            if fa.shape != (self.n_ue,):
                raise ValueError("need one fixed action per UE")
            keep.append(fa)
            pol.host_fixed_action = fa.ctypes.data
        T = int(T)
        if out is None:
            o, t = self._outputs(T, obs=obs, info=info)
            t['_struct'] = o
            if return_actions:
                t['actions'] = self._empty((T, self.num_envs, self.n_ue), torch.int32)
        else:
            o, t = out['_struct'], out
        aptr = ctypes.c_void_p(t['actions'].data_ptr()) if 'actions' in t else None
        check(self._L.dcb_rollout(self._h, ctypes.byref(pol), T, aptr, ctypes.byref(o), self._stream()))
        return t

    # ------------------------------------------------------------------ host-buffer path (e2e)
    def pinned_buffers(self):
        """Pinned host buffers for step_host: actions int32 [K,N], obs, reward, lost_conn."""
        if self._pinned is None:
            K, N = self.num_envs, self.n_ue
            self._pinned = dict(
                actions=torch.empty((K, N), dtype=torch.int32).pin_memory(),
                obs=torch.empty((K,) + self.obs_shape, dtype=torch.float32).pin_memory(),
                reward=torch.empty((K,) + self.reward_shape, dtype=torch.float32).pin_memory(),
                lost_conn=torch.empty((K, N), dtype=torch.uint8).pin_memory())
        return self._pinned

    def step_host(self, actions=None):
        """
        One step through host memory: H2D copy of the actions, the step, D2H copies of obs / reward / lost_conn, and a
        stream synchronise -- all inside the C-ABI call dcb_step_host.  `actions`: None (already written into
        pinned_buffers()['actions']) or an int array [K, N].
        """
        pb = self.pinned_buffers()
        if actions is not None:
            pb['actions'].copy_(torch.as_tensor(np.asarray(actions, dtype=np.int32)))
        if self._dynamic:
            # arrivals / departures edit the action buffer on the device (base.py:429-443): host -> device, events + step,
            # device -> the same pinned buffers
            obs, rew, _, info = self.step(pb['actions'].to(self.device, non_blocking=True), info=False)
            pb['obs'].copy_(obs, non_blocking=True)
            pb['reward'].copy_(rew, non_blocking=True)
            pb['lost_conn'].copy_(info['lost_conn'], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            return pb['obs'], pb['reward'], None, {'lost_conn': pb['lost_conn']}
        self._before_steps(1)
        self._t += 1
        check(self._L.dcb_step_host(self._h, ctypes.c_void_p(pb['actions'].data_ptr()),
                                    ctypes.c_void_p(pb['obs'].data_ptr()), ctypes.c_void_p(pb['reward'].data_ptr()),
                                    ctypes.c_void_p(pb['lost_conn'].data_ptr()), self._stream()))
        return pb['obs'], pb['reward'], None, {'lost_conn': pb['lost_conn']}

    def pinned_fragment_buffers(self, T):
        """Pinned host buffers for step_many_host: actions int32 [T,K,N], obs [T,K,...], reward [T,K,...], lost_conn."""
        K, N = self.num_envs, self.n_ue
        return dict(
            actions=torch.empty((T, K, N), dtype=torch.int32).pin_memory(),
            obs=torch.empty((T, K) + self.obs_shape, dtype=torch.float32).pin_memory(),
            reward=torch.empty((T, K) + self.reward_shape, dtype=torch.float32).pin_memory(),
            lost_conn=torch.empty((T, K, N), dtype=torch.uint8).pin_memory())

    def step_many_host(self, bufs, T=None, chunk_steps=0, actions=None):
        """
        T steps through HOST memory in one C-ABI call (dcb_step_many_host): `bufs` = pinned_fragment_buffers(T) with
        bufs['actions'] filled in (or `actions`: any pinned int32 host tensor [T, K, N], e.g. a slice of a larger action
        log); the steps run in chunks whose device -> host copies overlap the next chunk's compute, one synchronise at the
        end.  Returns (obs, reward, None, {'lost_conn'}) as [T, ...] host tensors.
        """
        if self._dynamic:
            raise NotImplementedError("step_many_host with a variable UE population (use step_many with device tensors)")
        acts = bufs['actions'] if actions is None else actions
        if not (acts.dtype == torch.int32 and acts.is_contiguous() and not acts.is_cuda):
            raise ValueError("actions must be a contiguous int32 host tensor [T, K, N]")
        T = int(acts.shape[0]) if T is None else int(T)
        self._before_steps(T)
        self._t += T
        check(self._L.dcb_step_many_host(self._h, ctypes.c_void_p(acts.data_ptr()), T,
                                         ctypes.c_void_p(bufs['obs'].data_ptr()),
                                         ctypes.c_void_p(bufs['reward'].data_ptr()),
                                         ctypes.c_void_p(bufs['lost_conn'].data_ptr()), int(chunk_steps), self._stream()))
        return bufs['obs'][:T], bufs['reward'][:T], None, {'lost_conn': bufs['lost_conn'][:T]}

    # ------------------------------------------------------------------ state snapshots
    def get_state(self):
        K, N = self.num_envs, self.n_ue
        st = dict(pos=np.zeros((K, N, 2)), mask=np.zeros((K, N), dtype=np.uint64), ewma=np.zeros((K, N)),
                  movement=np.zeros((K, N, 5)), time=np.zeros(K, dtype=np.int32))
        s = DcbStateHost(**{k: v.ctypes.data for k, v in st.items()})
        check(self._L.dcb_get_state(self._h, ctypes.byref(s)))
        return st

    def set_state(self, **arrays):
        """Inject any of pos [K,N,2] f64, mask [K,N] u64, ewma [K,N] f64, movement [K,N,5] f64, time [K] i32."""
        K, N = self.num_envs, self.n_ue
        spec = dict(pos=((K, N, 2), np.float64), mask=((K, N), np.uint64), ewma=((K, N), np.float64),
                    movement=((K, N, 5), np.float64), time=((K,), np.int32))
        keep = {}
        s = DcbStateHost()
        for k, v in arrays.items():
            shape, dt = spec[k]
            a = np.ascontiguousarray(np.asarray(v, dtype=dt))
            if a.shape != shape:
                raise ValueError(f"{k} must have shape {shape}")
            keep[k] = a
            setattr(s, k, a.ctypes.data)
        check(self._L.dcb_set_state(self._h, ctypes.byref(s)))

    def mask_matrix(self, mask=None):
        """uint64 bitmask [K,N] -> uint8 [K,N,M] connection matrix"""
        if mask is None:
            mask = self.get_state()['mask']
        b = np.arange(self.n_bs, dtype=np.uint64)
        return ((mask[..., None] >> b) & np.uint64(1)).astype(np.uint8)
