"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (deepcomp_b200/).

Imports the reference's env code *unmodified* from /root/reference under the stubs of
``oracle/ref_stubs.py`` and drives it in lock-step, recording a per-step trace.  Works only
where /root/reference exists (this container); the GPU box uses the committed fixtures in
``tests/golden/`` that ``oracle/make_golden.py`` produced with this module.

Scenario construction mirrors what the reference's factory does without importing it
(deepcomp/util/env_setup.py needs real RLlib): ``env_config`` keys as in env_setup.py:247-256,
UE creation as in env_setup.py:145-161, sharing mix as in env_setup.py:40-49.
"""
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get('DEEPCOMP_REFERENCE', '/root/reference')

SHARING_MIX = ['resource-fair', 'rate-fair', 'proportional-fair']      # env_setup.py:48


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'deepcomp', 'env'))


def _import_reference():
    from . import ref_stubs
    ref_stubs.install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from deepcomp.env.entities.map import Map
    from deepcomp.env.entities.station import Basestation
    from deepcomp.env.entities.user import User
    from deepcomp.env.util.movement import RandomWaypoint, UniformMovement
    from deepcomp.env.multi_ue.central import CentralRelNormEnv, CentralMaxNormEnv, CentralNormDrEnv, CentralDrEnv
    from deepcomp.env.multi_ue.multi_agent import MultiAgentMobileEnv, SeqMultiAgentMobileEnv
    from deepcomp.env.single_ue.variants import MaxNormEnv
    from shapely.geometry import Point
    return dict(Map=Map, Basestation=Basestation, User=User, RandomWaypoint=RandomWaypoint,
                CentralRelNormEnv=CentralRelNormEnv, MultiAgentMobileEnv=MultiAgentMobileEnv, Point=Point,
                CentralMaxNormEnv=CentralMaxNormEnv, MaxNormEnv=MaxNormEnv, CentralNormDrEnv=CentralNormDrEnv,
                CentralDrEnv=CentralDrEnv, UniformMovement=UniformMovement,
                SeqMultiAgentMobileEnv=SeqMultiAgentMobileEnv)


def sharing_for_bs(sharing, b):
    """env_setup.py:40-49"""
    return sharing if sharing != 'mixed' else SHARING_MIX[b % 3]


def grid_layout(n_bs, pitch=100, border=10):
    """Synthetic BS layout of SURVEY.md section 8d: square grid, 100 m pitch, 10 m border."""
    cols = int(np.ceil(np.sqrt(n_bs)))
    rows = int(np.ceil(n_bs / cols))
    width = max(pitch * (cols - 1) + 2 * border, 120)
    height = max(pitch * (rows - 1) + 2 * border, 120)
    bs_xy = [(border + pitch * (b % cols), border + pitch * (b // cols)) for b in range(n_bs)]
    return width, height, bs_xy


def build_env(kind, n_ue, seed, bs_xy, map_wh, sharing='mixed', velocities='slow', reward='avg',
              episode_length=100, rand_episodes=False, init_pos=None, max_ues=None, ue_arrival=None,
              new_ue_interval=None, util_func='log', obs_norm='rel', obs_variant=None, obs_opts=None,
              uniform_moves=None, sequential=False):
    """
    Build a reference env.

    :param obs_norm: 'rel' (RelNormEnv observation) or 'max' (MaxNormEnv, variants.py:308-332): central =
        CentralMaxNormEnv (central.py:155-164); multi = MultiAgentMobileEnv composed with MaxNormEnv the way the
        reference composes its central variant (class X(MultiAgentMobileEnv, MaxNormEnv): get_ue_obs resolves to
        MaxNormEnv's, everything else to MultiAgentMobileEnv's; both classes unmodified)

    :param kind: 'central' (CentralRelNormEnv) or 'multi' (MultiAgentMobileEnv)
    :param velocities: 'slow' | 'fast' | number, or a list of those per UE
    :param init_pos: None (all 'random') or list of (x, y) per UE with numbers or 'random'
    :param max_ues, ue_arrival, new_ue_interval: variable UE population (base.py:80-84,433-443; env_setup.py:205-226)
    """
    R = _import_reference()
    m = R['Map'](width=map_wh[0], height=map_wh[1])
    bs_list = [R['Basestation'](chr(ord('A') + b) if len(bs_xy) <= 26 else f'BS{b}', R['Point'](x, y),
                                sharing_for_bs(sharing, b) if isinstance(sharing, str) else sharing[b])
               for b, (x, y) in enumerate(bs_xy)]
    if not isinstance(velocities, (list, tuple)):
        velocities = [velocities] * n_ue
    ue_list = []
    for i in range(n_ue):
        px, py = ('random', 'random') if init_pos is None else init_pos[i]
        if uniform_moves is not None and uniform_moves[i] is not None:
            mv = R['UniformMovement'](m, move_x=uniform_moves[i][0], move_y=uniform_moves[i][1])   # movement.py:26-80
        else:
            mv = R['RandomWaypoint'](m, velocity=velocities[i])
        ue_list.append(R['User'](str(i + 1), m, pos_x=px, pos_y=py, movement=mv, util_func=util_func))
    env_config = {
        'episode_length': episode_length, 'seed': seed, 'map': m, 'bs_list': bs_list, 'ue_list': ue_list,
        'rand_episodes': rand_episodes, 'new_ue_interval': new_ue_interval, 'reward': reward, 'max_ues': max_ues,
        'ue_arrival': None if ue_arrival is None else {int(t): int(n) for t, n in ue_arrival.items()},
        'log_metrics': True, 'dashboard': False, 'ue_details': False,
    }
    if sequential:
        assert kind == 'multi' and obs_variant is None and obs_norm == 'rel'
        return R['SeqMultiAgentMobileEnv'](env_config)           # multi_agent.py:110-179
    if obs_variant is not None:
        # CentralNormDrEnv (central.py:107-140) / CentralDrEnv (central.py:75-104, options read from env_config,
        # variants.py:68-73); the reference has no multi-agent class with these observations
        assert kind == 'central' and obs_variant in ('normdr', 'datarate')
        if obs_variant == 'datarate':
            opts = dict(dr_cutoff='auto', sub_req_dr=True, curr_dr_obs=False, ues_at_bs_obs=False, dist_obs=False,
                        next_dist_obs=False)
            opts.update(obs_opts or {})
            env_config.update(opts)
        cls = R['CentralNormDrEnv'] if obs_variant == 'normdr' else R['CentralDrEnv']
    elif obs_norm == 'max':
        cls = R['CentralMaxNormEnv'] if kind == 'central' else \
            type('MultiAgentMaxNormEnv', (R['MultiAgentMobileEnv'], R['MaxNormEnv']), {})
    else:
        cls = R['CentralRelNormEnv'] if kind == 'central' else R['MultiAgentMobileEnv']
    return cls(env_config)


class RefTrace:
    """Drive a reference env and record everything the parity tests compare."""

    def __init__(self, env, kind):
        self.env, self.kind = env, kind
        self._lost = None
        orig_move = env.move_ues

        def move_and_record():
            lost = orig_move()                      # base.py:337-348 (return value discarded by step, base.py:447)
            self._lost = [lost[ue] for ue in env.ue_list]
            return lost
        env.move_ues = move_and_record

    # ---- state snapshots -------------------------------------------------
    def positions(self):
        return np.array([[ue.pos.x, ue.pos.y] for ue in self.env.ue_list], dtype=np.float64)

    def mask(self):
        e = self.env
        return np.array([[int(bs in ue.bs_dr) for bs in e.bs_list] for ue in e.ue_list], dtype=np.uint8)

    def link_rates(self):
        e = self.env
        return np.array([[float(ue.bs_dr.get(bs, 0.0)) for bs in e.bs_list] for ue in e.ue_list], dtype=np.float64)

    def snr(self):
        e = self.env
        return np.array([[float(bs.snr(ue.pos)) for bs in e.bs_list] for ue in e.ue_list], dtype=np.float64)

    def curr_dr(self):
        return np.array([float(ue.curr_dr) for ue in self.env.ue_list], dtype=np.float64)

    def ewma(self):
        return np.array([float(ue.ewma_dr) for ue in self.env.ue_list], dtype=np.float64)

    def utility(self):
        return np.array([float(ue.utility) for ue in self.env.ue_list], dtype=np.float64)

    def movement(self):
        """velocity, waypoint x, waypoint y, pausing, curr_pause per UE"""
        return np.array([[ue.movement.move_x, ue.movement.move_y, -1.0, 0.0, 0.0] if hasattr(ue.movement, 'move_x') else
                         [ue.movement.velocity, ue.movement.waypoint.x, ue.movement.waypoint.y,
                          float(ue.movement.pausing), ue.movement.curr_pause] for ue in self.env.ue_list],
                        dtype=np.float64)

    # ---- obs / reward flattening (RLlib Dict-flattening order = sorted keys) -------------
    def _sequential(self):
        return hasattr(self.env, 'ue_order_idx')

    def flat_obs(self, obs):
        e = self.env
        if self._sequential():                  # SeqMultiAgentMobileEnv: the current UE's observation only
            (o,) = obs.values()
            return np.concatenate([np.asarray(o[k], dtype=np.float64).ravel() for k in sorted(o.keys())])
        if self.kind == 'central':
            return np.concatenate([np.asarray(obs[k], dtype=np.float64) for k in sorted(obs.keys())])
        rows = []
        for ue in e.ue_list:
            o = obs[ue.id]
            rows.append(np.concatenate([np.asarray(o[k], dtype=np.float64).ravel() for k in sorted(o.keys())]))
        return self._pad(np.stack(rows))

    def flat_reward(self, reward):
        if self._sequential():
            (r,) = reward.values()
            return np.float64(r)
        if self.kind == 'central':
            return np.float64(reward)
        return self._pad(np.array([float(reward[ue.id]) for ue in self.env.ue_list], dtype=np.float64))

    def to_action(self, a):
        """a: int array [N] -> the action object the env class expects"""
        if self._sequential():                  # only the current UE acts (multi_agent.py:21-30 skips the others)
            return {self.env.curr_ue.id: int(a[self.env.ue_order_idx])}
        if self.kind == 'central':
            return np.asarray(a, dtype=np.int64)          # length max_ues; entries beyond the UEs present are ignored
        return {ue.id: int(a[i]) for i, ue in enumerate(self.env.ue_list)}

    def _pad(self, a):
        """rows of the UEs present (list order) first, zero rows up to max_ues (variable population)"""
        a = np.asarray(a)
        missing = self.env.max_ues - a.shape[0]
        if missing <= 0:
            return a
        return np.concatenate([a, np.zeros((missing,) + a.shape[1:], dtype=a.dtype)])

    def snapshot(self):
        d = dict(pos=self.positions(), mask=self.mask(), link_rates=self.link_rates(), snr=self.snr(),
                 curr_dr=self.curr_dr(), ewma=self.ewma(), utility=self.utility(), movement=self.movement())
        d = {k: self._pad(v) for k, v in d.items()}
        d['num_ue'] = len(self.env.ue_list)
        return d

    def reset(self):
        obs = self.env.reset()
        out = self.snapshot()
        out['obs'] = self.flat_obs(obs)
        return out

    def step(self, a):
        self._lost = None
        obs, reward, done, info = self.env.step(self.to_action(a))
        if self._lost is None:                  # a sequential sub-step without a move
            self._lost = [0] * len(self.env.ue_list)
        self.last_obs = obs
        out = self.snapshot()
        out['obs'] = self.flat_obs(obs)
        out['reward'] = self.flat_reward(reward)
        out['lost_conn'] = self._pad(np.array(self._lost, dtype=np.int32))
        out['done'] = done
        out['info'] = info
        if self._sequential():
            # multi_agent.py:143-146 wraps MultiAgentMobileEnv.info() (already a dict per UE id) once more
            (inner,) = info.values()
            info0 = next(iter(inner.values()))
        elif self.kind == 'multi':
            info0 = info[self.env.ue_list[0].id]
        else:
            info0 = info
        out['sum_utility'] = np.float64(info0['scalar_metrics']['sum_utility'])
        out['time'] = info0['time']
        return out
