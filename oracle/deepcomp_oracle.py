"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's env.step hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module; the product path (``deepcomp_b200/``) never does and fails loudly
when its CUDA library is missing.

Parity status: **pinned against the live reference** -- ``tests/test_oracle_vs_reference.py`` runs
this restatement in lock-step with the unmodified reference env imported from /root/reference
(under ``oracle/ref_stubs.py``) and asserts *bit-identical* positions, masks, lost-connection
counts, link rates, utilities, observations and rewards; ``oracle/make_golden.py`` freezes traces of
the reference itself into ``tests/golden/*.npz`` so that the same pin travels to the GPU box.  The
reference has no tests / golden vectors of its own; the one arithmetic piece
that lives in an un-vendored dependency is shapely==1.7.0 (setup.py:17) ``Point.distance`` ==
GEOS ``sqrt(dx*dx+dy*dy)``, restated in ``_dist``.

The restatement keeps the reference's per-object evaluation order (dict insertion order of
``ue.bs_dr``, list order of ``bs.conn_ues``) so that floating-point sums associate identically.
All file:line citations are relative to /root/reference/deepcomp/.

Beyond what the CUDA path offers today the oracle also restates (pinned on traces of the reference, kernels to follow):
CentralNormDrEnv / CentralDrEnv observations (``obs_variant``), UniformMovement (``uniform_moves``) and
SeqMultiAgentMobileEnv (``sequential``).
"""
import math
import random
from fractions import Fraction

import numpy as np

# util/constants.py:28,34-35,40-41
EPSILON = 1e-16
FAIR_WEIGHT_ALPHA = 1
FAIR_WEIGHT_BETA = 1
MIN_UTILITY = -20
MAX_UTILITY = 20
# env/entities/station.py:10
SNR_THRESHOLD = 2e-8
MAX_SNR_THRESHOLD = 7e-6                    # MaxNormEnv.MAX_SNR_THRESHOLD (single_ue/variants.py:311)
# env/entities/station.py:26-30
BW = 9e6
FREQUENCY = 2500
NOISE = 1e-9
TX_POWER = 30
BS_HEIGHT = 50
UE_HEIGHT = 1.5

SHARING_MIX = ['resource-fair', 'rate-fair', 'proportional-fair']          # util/env_setup.py:48
SHARING_CODE = {'resource-fair': 0, 'rate-fair': 1, 'proportional-fair': 2, 'max-cap': 3}


def sharing_for_bs(sharing, b):
    """util/env_setup.py:40-49"""
    return sharing if sharing != 'mixed' else SHARING_MIX[b % 3]


def grid_layout(n_bs, pitch=100, border=10):
    """Synthetic layout of SURVEY.md section 8d (pitch = cli.py:41 --bs-dist default, border = env_setup.py:87)."""
    cols = int(np.ceil(np.sqrt(n_bs)))
    rows = int(np.ceil(n_bs / cols))
    width = max(pitch * (cols - 1) + 2 * border, 120)
    height = max(pitch * (rows - 1) + 2 * border, 120)
    bs_xy = [(border + pitch * (b % cols), border + pitch * (b // cols)) for b in range(n_bs)]
    return width, height, bs_xy


def path_loss_consts():
    """env/entities/station.py:110-116 -- evaluated exactly as the reference does (np.log10 on scalars)."""
    ch = 0.8 + (1.1 * np.log10(FREQUENCY) - 0.7) * UE_HEIGHT - 1.56 * np.log10(FREQUENCY)
    const1 = 69.55 + 26.16 * np.log10(FREQUENCY) - 13.82 * np.log10(BS_HEIGHT) - ch
    const2 = 44.9 - 6.55 * np.log10(BS_HEIGHT)
    return const1, const2


_CONST1, _CONST2 = path_loss_consts()


def _dist(ax, ay, bx, by):
    """shapely 1.7.0 Point.distance -> GEOS: sqrt(dx*dx + dy*dy) in double (station.py:124, movement.py:142)."""
    dx = ax - bx
    dy = ay - by
    return math.sqrt(dx * dx + dy * dy)


def _fma(a, b, c):
    """Correctly rounded a*b+c (Python 3.12 has no math.fma)."""
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def _norm2(vx, vy):
    """
    np.linalg.norm of a 2-vector (movement.py:151) == sqrt(x.dot(x)); the dot product goes to OpenBLAS ddot whose
    x86 FMA3 kernels accumulate acc = fma(x_i, x_i, acc) -- i.e. sqrt(fma(vy, vy, vx*vx)).  Verified bit-identical
    to np.linalg.norm on 1e5 random vectors in this container ("pinned arithmetic").
    """
    return math.sqrt(_fma(vy, vy, vx * vx))


def snr_of_distance(distance):
    """station.py:110-127: Okumura-Hata path loss -> received power -> SNR"""
    pl = _CONST1 + _CONST2 * np.log10(distance + EPSILON)
    signal = 10 ** ((TX_POWER - pl) / 10)
    return signal / NOISE


def log_utility(curr_dr):
    """env/util/utility.py:36-54"""
    if curr_dr == 0:
        return MIN_UTILITY
    return np.clip(10 * np.log10(curr_dr), MIN_UTILITY, MAX_UTILITY)


def threshold_distance():
    """Largest double d with snr(d) > SNR_THRESHOLD (station.py:224); used by fast restatements (C / CUDA)."""
    lo, hi = 60.0, 80.0
    assert snr_of_distance(lo) > SNR_THRESHOLD >= snr_of_distance(hi)
    while True:
        mid = 0.5 * (lo + hi)
        if mid == lo or mid == hi:
            break
        if snr_of_distance(mid) > SNR_THRESHOLD:
            lo = mid
        else:
            hi = mid
    return lo


class _UE:
    __slots__ = ('idx', 'id', 'x', 'y', 'init_x', 'init_y', 'init_velocity', 'velocity', 'wx', 'wy', 'pausing',
                 'curr_pause', 'rng', 'mrng', 'bs_dr', 'ewma_dr', 'uniform', 'move_x', 'move_y')


class OracleEnv:
    """
    One env instance, restating MobileEnv (+ CentralRelNormEnv / MultiAgentMobileEnv) for the in-scope feature set:
    RandomWaypoint movement, log utility, the four sharing models, fixed or variable UE population (`max_ues`,
    `ue_arrival`, `new_ue_interval`: base.py:80-84, 433-443, 592-617; per-UE arrays are padded to max_ues).

    Variable population: as in the reference, reset() re-seeds the UEs of the *current* list by their current list
    position before it restores the original list (base.py:132-143, 169-189) -- after departures an original UE can
    come back with another UE's seed, and an original UE that was removed keeps drawing from its old streams.

    kind: 'central' (multi_ue/central.py:143-152) or 'multi' (multi_ue/multi_agent.py:6-107)
    """

    def __init__(self, kind, n_ue, bs_xy, map_wh, sharing='mixed', velocities='slow', seed=None, reward='avg',
                 episode_length=100, rand_episodes=False, init_pos=None, pause_duration=2, border_buffer=10,
                 max_ues=None, ue_arrival=None, new_ue_interval=None, util_func='log', dr_req=1, obs_norm='rel',
                 obs_variant=None, obs_opts=None, uniform_moves=None, sequential=False):
        assert kind in ('central', 'multi')
        # sequential: SeqMultiAgentMobileEnv (multi_ue/multi_agent.py:110-179; restated for the next round, no CUDA path
        # yet): one UE acts per call, the UEs move and time advances after the last one; obs / reward of the next UE only
        assert not sequential or kind == 'multi'
        self.sequential = sequential
        self.ue_order_idx = 0                                   # multi_agent.py:119 (never reset, not even by reset())
        # uniform_moves: None, or per UE None (RandomWaypoint) / (move_x, move_y) with numbers or 'slow' / 'fast' =
        # UniformMovement (util/movement.py:26-80; restated for the next round's kernels, no CUDA path yet)
        self.uniform_moves = uniform_moves
        assert obs_norm in ('rel', 'max')                       # RelNormEnv / MaxNormEnv (variants.py:271-303 / 308-332)
        self.obs_norm = obs_norm
        # other observation classes of the reference (restated for the next round's kernels; no CUDA path yet):
        # 'normdr' = NormDrMobileEnv / CentralNormDrEnv (variants.py:173-250, central.py:107-140),
        # 'datarate' = DatarateMobileEnv / CentralDrEnv (variants.py:42-170, central.py:75-104) with its env_config options
        assert obs_variant in (None, 'normdr', 'datarate')
        self.obs_variant = obs_variant
        self.obs_opts = dict(dr_cutoff='auto', sub_req_dr=True, curr_dr_obs=False, ues_at_bs_obs=False, dist_obs=False,
                             next_dist_obs=False)
        self.obs_opts.update(obs_opts or {})
        if obs_variant == 'datarate':                           # variants.py:75-79
            o = self.obs_opts
            assert not (o['dr_cutoff'] == 'auto' and not o['sub_req_dr'])
            assert (not o['curr_dr_obs']) or (o['dr_cutoff'] == 'auto' and o['sub_req_dr'])
            assert o['dist_obs'] or not o['next_dist_obs']
        assert util_func in ('log', 'step')                     # 'linear' fails the reference's own assert (utility.py:18)
        self.util_func, self.dr_req = util_func, dr_req
        self.kind = kind
        self.n_ue = n_ue
        # variable population (base.py:80-84): per-UE arrays have max_ues rows, the UEs present come first
        self.max_ues = n_ue if max_ues is None else int(max_ues)
        assert self.max_ues >= n_ue
        self.ue_arrival = None if ue_arrival is None else {int(t): int(n) for t, n in ue_arrival.items()}
        self.new_ue_interval = new_ue_interval
        self.map_rng = random.Random()                          # entities/map.py:30
        self.glob_rng = random.Random()                         # the `random` module itself (base.py:134, 612)
        self.bs_xy = [(float(x), float(y)) for x, y in bs_xy]
        self.n_bs = len(bs_xy)
        # entities/map.py:20-21
        self.width, self.height = int(map_wh[0]), int(map_wh[1])
        if isinstance(sharing, str):
            sharing = [sharing_for_bs(sharing, b) for b in range(self.n_bs)]
        self.sharing = list(sharing)
        if not isinstance(velocities, (list, tuple)):
            velocities = [velocities] * n_ue
        self.reward_agg = reward
        self.episode_length = episode_length
        self.rand_episodes = rand_episodes
        self.pause_duration = pause_duration
        self.border_buffer = border_buffer
        self.env_seed = seed
        self.time = 0
        self.total_utility = 0
        self.ues = []
        for i in range(n_ue):
            ue = _UE()
            ue.idx = i
            ue.id = str(i + 1)                                  # util/env_setup.py:148-160
            ue.init_x, ue.init_y = ('random', 'random') if init_pos is None else init_pos[i]
            ue.init_velocity = velocities[i]
            ue.uniform = None if uniform_moves is None else uniform_moves[i]
            ue.rng = random.Random()                            # entities/user.py:38
            ue.mrng = random.Random()                           # util/movement.py:14
            ue.bs_dr = {}
            ue.ewma_dr = 0
            ue.x = ue.y = 0.0
            ue.pausing, ue.curr_pause = False, 0
            self.ues.append(ue)
        self.original_ues = list(self.ues)                      # base.py:52 original_ue_list
        # per-BS connected-UE lists (station.py:18), in connection order
        self.conn_ues = [[] for _ in range(self.n_bs)]
        self.seed(seed)
        self.last_lost_conn = [0] * n_ue

    # ------------------------------------------------------------------ seeding / reset
    def seed(self, seed=None):
        """single_ue/base.py:132-143 (+ user.py:94-96): UE i (1-based) gets seed+100*i for BOTH of its RNGs"""
        if seed is not None:
            self.glob_rng.seed(seed)                            # base.py:134 random.seed(seed)
            self.map_rng.seed(seed)                             # base.py:136 -> map.py:49-50
            offset = 0
            for ue in self.ues:
                offset += 100
                ue.rng.seed(seed + offset)
                ue.mrng.seed(seed + offset)

    def _movement_reset(self, ue):
        """util/movement.py:110-130 (RandomWaypoint.reset); :47-64 (UniformMovement.reset)"""
        if getattr(ue, 'uniform', None) is not None:
            mv = []
            for init in ue.uniform:                             # move_x first, then move_y
                if init == 'slow':
                    mv.append(ue.mrng.randint(1, 5))
                elif init == 'fast':
                    mv.append(ue.mrng.randint(10, 20))
                else:
                    mv.append(init)
            ue.move_x, ue.move_y = mv
            ue.velocity, ue.wx, ue.wy, ue.pausing, ue.curr_pause = 0, 0.0, 0.0, False, 0
            return
        if ue.init_velocity == 'slow':
            ue.velocity = ue.mrng.randint(1, 3)
        elif ue.init_velocity == 'fast':
            ue.velocity = ue.mrng.randint(5, 10)
        else:
            ue.velocity = ue.init_velocity
        x = ue.mrng.randint(self.border_buffer, int(self.width - self.border_buffer))
        y = ue.mrng.randint(self.border_buffer, int(self.height - self.border_buffer))
        ue.wx, ue.wy = float(x), float(y)
        ue.pausing = False
        ue.curr_pause = 0

    def reset(self):
        """single_ue/base.py:169-189; user.py:98-116; station.py:106-108"""
        if not self.rand_episodes:
            self.seed(self.env_seed)                            # base.py:171-173: seeds the CURRENT list by position ...
        self.ues = list(self.original_ues)                      # base.py:176-182: ... then restores the original list
        self.time = 0
        for ue in self.ues:
            px = ue.init_x
            if px == 'random':
                px = ue.rng.randint(0, int(self.width))
            py = ue.init_y
            if py == 'random':
                py = ue.rng.randint(0, int(self.height))
            ue.x, ue.y = float(px), float(py)
            self._movement_reset(ue)
            ue.bs_dr = {}
            ue.ewma_dr = 0
        self.conn_ues = [[] for _ in range(self.n_bs)]
        return self.get_obs()

    # ------------------------------------------------------------------ radio model
    def snr(self, b, ue):
        """station.py:122-127"""
        bx, by = self.bs_xy[b]
        return snr_of_distance(_dist(bx, by, ue.x, ue.y))

    def can_connect(self, b, ue):
        """station.py:222-226"""
        return self.snr(b, ue) > SNR_THRESHOLD

    def data_rate_unshared(self, b, ue):
        """station.py:129-138"""
        return BW * np.log2(1 + self.snr(b, ue))

    def priority(self, b, ue):
        """station.py:140-150"""
        return (self.data_rate_unshared(b, ue) ** FAIR_WEIGHT_ALPHA) / (ue.ewma_dr ** FAIR_WEIGHT_BETA + EPSILON)

    def data_rate_shared(self, b, ue, dr_ue_unshared):
        """station.py:152-202"""
        conn = self.conn_ues[b]
        already = ue in conn
        if not already:
            ue.bs_dr[b] = self.data_rate_unshared(b, ue)
            conn.append(ue)
        model = self.sharing[b]
        if model == 'resource-fair':
            dr = dr_ue_unshared / len(conn)
        elif model == 'rate-fair':
            total_inverse_dr = sum([1 / self.data_rate_unshared(b, o) for o in conn])
            dr = 1 / total_inverse_dr
        elif model == 'max-cap':
            max_ue_idx = np.argmax([self.data_rate_unshared(b, o) for o in conn])
            dr = 0
            if conn.index(ue) == max_ue_idx:
                dr = self.data_rate_unshared(b, ue)
        elif model == 'proportional-fair':
            frac = self.priority(b, ue) / (sum([self.priority(b, o) for o in conn]) + EPSILON)
            dr = frac * dr_ue_unshared
        else:
            raise AssertionError(model)
        if not already:
            del ue.bs_dr[b]
            conn.remove(ue)
        return dr

    def data_rate(self, b, ue):
        """station.py:204-220"""
        if not self.can_connect(b, ue):
            return 0
        return self.data_rate_shared(b, ue, self.data_rate_unshared(b, ue))

    # ------------------------------------------------------------------ UE
    @staticmethod
    def curr_dr(ue):
        """user.py:64-69"""
        return sum(list(ue.bs_dr.values()))

    def utility(self, ue):
        """user.py:76-92; env/util/utility.py:23-33 (step), 36-54 (log)"""
        if self.util_func == 'step':
            return MAX_UTILITY if self.curr_dr(ue) >= self.dr_req else MIN_UTILITY
        return log_utility(self.curr_dr(ue))

    def connect_to_bs(self, ue, b):
        """user.py:190-229 with disconnect=True (the only way the env calls it, base.py:263)"""
        if b in ue.bs_dr:
            del ue.bs_dr[b]
            self.conn_ues[b].remove(ue)
            return
        if self.can_connect(b, ue):
            ue.bs_dr[b] = self.data_rate(b, ue)
            self.conn_ues[b].append(ue)

    def movement_step(self, ue):
        """util/movement.py:132-181 (RandomWaypoint.step); :66-80 (UniformMovement.step)"""
        if getattr(ue, 'uniform', None) is not None:
            nx, ny = ue.x + ue.move_x, ue.y + ue.move_y
            # Point.within(map.shape): strictly inside the rectangle (a point on the border is not within)
            if not (0 < nx < self.width and 0 < ny < self.height):
                ue.move_x, ue.move_y = -ue.move_x, -ue.move_y   # bounce: BOTH components flip, no second check
                nx, ny = ue.x + ue.move_x, ue.y + ue.move_y
            ue.x, ue.y = float(nx), float(ny)
            return
        if ue.x == ue.wx and ue.y == ue.wy:
            ue.pausing = True
        if ue.pausing:
            if ue.curr_pause < self.pause_duration:
                ue.curr_pause += 1
                return
            self._movement_reset(ue)
        ue.x, ue.y = self.step_towards_waypoint(ue)

    @staticmethod
    def step_towards_waypoint(ue):
        """util/movement.py:132-156: the position one step closer to the waypoint (the UE itself is not moved)"""
        if _dist(ue.x, ue.y, ue.wx, ue.wy) <= ue.velocity:
            return ue.wx, ue.wy
        vx = ue.wx - ue.x
        vy = ue.wy - ue.y
        norm = _norm2(vx, vy)
        return ue.x + ue.velocity * (vx / norm), ue.y + ue.velocity * (vy / norm)

    def move(self, ue, weight=0.9):
        """user.py:148-188"""
        self.movement_step(ue)
        remove = [b for b in ue.bs_dr if not self.can_connect(b, ue)]
        for b in remove:
            del ue.bs_dr[b]
            self.conn_ues[b].remove(ue)
        ue.ewma_dr = weight * self.curr_dr(ue) + (1 - weight) * ue.ewma_dr
        return len(remove)

    # ------------------------------------------------------------------ env
    @staticmethod
    def calc_reward(utility, penalty):
        """single_ue/base.py:158-167"""
        clip_util = np.clip(utility, MIN_UTILITY, MAX_UTILITY)
        return np.clip(clip_util + penalty, MIN_UTILITY, MAX_UTILITY) / MAX_UTILITY

    def update_ue_drs_rewards(self, update_only=False):
        """single_ue/base.py:315-335 (penalties are identically 0, base.py:257)"""
        rewards = []
        for ue in self.ues:
            for b in ue.bs_dr:                                  # user.py:143-146
                ue.bs_dr[b] = self.data_rate(b, ue)
            rewards.append(0 if update_only else self.calc_reward(self.utility(ue), 0))
        return rewards

    def get_ue_obs_normdr(self, ue):
        """NormDrMobileEnv.get_ue_obs (single_ue/variants.py:198-250): the *shared* rate the UE gets or would get from every
        BS (station.py:204-220: 0 out of range; a UE that is not connected is counted in temporarily), cut at 100"""
        cutoff = 100
        bs_dr = [min(self.data_rate(b, ue), cutoff) / cutoff for b in range(self.n_bs)]
        bs_conn = [int(b in ue.bs_dr) for b in range(self.n_bs)]
        return {'dr': bs_dr, 'connected': bs_conn, 'dr_total': [min(self.curr_dr(ue), cutoff) / cutoff]}

    def get_ue_obs_datarate(self, ue):
        """DatarateMobileEnv.get_ue_obs (single_ue/variants.py:127-170)"""
        o, req = self.obs_opts, self.dr_req
        obs = {}
        if o['dr_cutoff'] == 'auto':
            obs['dr'] = [min(self.data_rate(b, ue) - req, req) / req for b in range(self.n_bs)]
        elif o['sub_req_dr']:
            obs['dr'] = [min(self.data_rate(b, ue) - req, o['dr_cutoff']) for b in range(self.n_bs)]
        else:
            obs['dr'] = [min(self.data_rate(b, ue), o['dr_cutoff']) for b in range(self.n_bs)]
        obs['connected'] = [int(b in ue.bs_dr) for b in range(self.n_bs)]
        if o['curr_dr_obs']:
            total = self.curr_dr(ue)
            total -= req
            total = min(total, req)
            obs['dr_total'] = [total / req]
        if o['ues_at_bs_obs']:
            obs['ues_at_bs'] = [len(self.conn_ues[b]) for b in range(self.n_bs)]
        diagonal = np.sqrt(self.width ** 2 + self.height ** 2)                           # entities/map.py:26
        if o['dist_obs']:
            obs['dist'] = [_dist(ue.x, ue.y, bx, by) / diagonal for bx, by in self.bs_xy]
        if o['next_dist_obs']:
            nx, ny = self.step_towards_waypoint(ue)
            obs['next_dist'] = [_dist(nx, ny, bx, by) / diagonal for bx, by in self.bs_xy]
        return obs

    def get_ue_obs(self, ue):
        """single_ue/variants.py:271-303"""
        if self.obs_variant == 'normdr':
            return self.get_ue_obs_normdr(ue)
        if self.obs_variant == 'datarate':
            return self.get_ue_obs_datarate(ue)
        bs_conn = [int(b in ue.bs_dr) for b in range(self.n_bs)]
        bs_dr = [self.snr(b, ue) for b in range(self.n_bs)]
        max_dr = max(bs_dr)
        if max_dr == 0:
            bs_norm_dr = [0 for _ in bs_dr]
        else:
            bs_norm_dr = [dr / max_dr for dr in bs_dr]
        if self.obs_norm == 'max':
            # MaxNormEnv.get_ue_obs (variants.py:319-332): cap at MAX_SNR_THRESHOLD, subtract the required SNR, normalise
            bs_norm_dr = [(min(self.snr(b, ue), MAX_SNR_THRESHOLD) - SNR_THRESHOLD) / (MAX_SNR_THRESHOLD - SNR_THRESHOLD)
                          for b in range(self.n_bs)]
        utility = [self.utility(ue) / MAX_UTILITY]
        ues_at_bs = [len(self.conn_ues[b]) / len(self.ues) for b in range(self.n_bs)]    # self.num_ue, variants.py:296
        avg_util = []
        for b in range(self.n_bs):                              # station.py:71-76
            c = self.conn_ues[b]
            avg_util.append((np.mean([self.utility(o) for o in c]) if len(c) > 0 else 0) / MAX_UTILITY)
        return {'connected': bs_conn, 'dr': bs_norm_dr, 'utility': utility, 'ues_at_bs': ues_at_bs,
                'util_at_bs': avg_util}

    def get_obs(self):
        """
        Flat observation in RLlib's Dict-flattening order (alphabetical keys).
        central (central.py:31-57,147-152): [connected(N*M) | dr(N*M) | utility(N)]
        multi (multi_agent.py:32-37, variants.py:255-269): per UE [connected(M) | dr(M) | ues_at_bs(M) | util_at_bs(M) | utility(1)]
        """
        per_ue = [self.get_ue_obs(ue) for ue in self.ues]
        missing = self.max_ues - len(self.ues)
        if self.kind == 'central':
            keys = ('connected', 'dr', 'utility')
            if self.obs_variant is not None:
                # CentralNormDrEnv / CentralDrEnv (central.py:75-140): the keys of the class's Dict space, sorted
                keys = sorted(per_ue[0].keys())
            out = []
            for key in keys:
                for o in per_ue:
                    out.extend(o[key])
                out.extend([0] * (missing * len(per_ue[0][key])))   # central.py:46-55: zeros for the UEs not there
            return np.asarray(out, dtype=np.float64)
        keys = ('connected', 'dr', 'ues_at_bs', 'util_at_bs', 'utility') if self.obs_variant is None \
            else sorted(per_ue[0].keys())
        rows = np.stack([np.concatenate([np.asarray(o[k], dtype=np.float64) for k in keys]) for o in per_ue])
        return self._pad(rows)

    def step_reward(self, rewards):
        if self.kind == 'central':
            # multi_ue/central.py:65-73
            if self.reward_agg == 'avg':
                return np.float64(np.mean(rewards))
            if self.reward_agg == 'sum':
                return np.float64(sum(rewards))
            if self.reward_agg == 'min':
                return np.float64(min(rewards))
            raise NotImplementedError(self.reward_agg)
        # multi_ue/multi_agent.py:39-95
        out = []
        for ue in self.ues:
            agg_util = self.utility(ue)
            in_range = [b for b in range(self.n_bs) if self.can_connect(b, ue)]
            if len(in_range) > 0:
                if self.reward_agg == 'avg':
                    num_neighbors = sum([len(self.conn_ues[b]) for b in in_range])
                    if num_neighbors > 0:
                        total = sum([sum([self.utility(o) for o in self.conn_ues[b]]) for b in in_range])
                        if len(ue.bs_dr) == 0:
                            agg_util = (total + self.utility(ue)) / (num_neighbors + 1)
                        else:
                            agg_util = total / num_neighbors
                elif self.reward_agg == 'sum':
                    # user.py:238-244 builds a *set* of neighbours; its iteration order is hash order in the
                    # reference -- summed in UE-index order here (differences are O(1 ulp))
                    where = {id(o): j for j, o in enumerate(self.ues)}
                    neigh = set()
                    for b in ue.bs_dr:
                        neigh.update(where[id(o)] for o in self.conn_ues[b])
                    agg_util = sum([rewards[j] for j in sorted(neigh)])
                elif self.reward_agg == 'min':
                    mins = []
                    for b in in_range:                          # station.py:78-83
                        c = self.conn_ues[b]
                        mins.append(min([self.utility(o) for o in c]) if len(c) > 0 else MAX_UTILITY)
                    agg_util = min(mins + [self.utility(ue)])
                else:
                    raise NotImplementedError(self.reward_agg)
            out.append(agg_util)
        return self._pad(np.asarray(out, dtype=np.float64))

    # ------------------------------------------------------------------ variable population
    def _pad(self, a):
        a = np.asarray(a)
        missing = self.max_ues - a.shape[0]
        if missing <= 0:
            return a
        return np.concatenate([a, np.zeros((missing,) + a.shape[1:], dtype=a.dtype)])

    def rand_border_point(self):
        """entities/map.py:52-65 (min_x = min_y = 0)"""
        x = self.map_rng.randint(0, self.width)
        y = self.map_rng.randint(0, self.height)
        border = self.map_rng.choice(['left', 'right', 'top', 'bottom'])
        return {'left': (0, y), 'right': (self.width - 1, y), 'top': (x, self.height - 1), 'bottom': (x, 0)}[border]

    def add_new_ue(self):
        """single_ue/base.py:592-608"""
        new_id = int(self.ues[-1].id) + 1
        px, py = self.rand_border_point()
        ue = _UE()
        ue.idx = None
        ue.id = str(new_id)
        ue.init_x, ue.init_y = px, py
        ue.init_velocity = 'slow'
        ue.rng = random.Random()
        ue.mrng = random.Random()
        seed = new_id * 100 if self.env_seed is None else self.env_seed + new_id * 100
        ue.rng.seed(seed)
        ue.mrng.seed(seed)
        ue.x, ue.y = float(px), float(py)                       # user.py:111-116 reset(): fixed position, then
        self._movement_reset(ue)                                # movement.reset()
        ue.bs_dr = {}
        ue.ewma_dr = 0
        self.ues.append(ue)

    def remove_ue(self):
        """single_ue/base.py:610-617: a uniformly random UE of the list, drawn from the global `random` module"""
        idx = self.glob_rng.randint(0, len(self.ues) - 1)
        ue = self.ues.pop(idx)
        for b in list(ue.bs_dr):                                # user.py:231-236 disconnect_from_all
            del ue.bs_dr[b]
            self.conn_ues[b].remove(ue)

    def step_sequential(self, actions):
        """SeqMultiAgentMobileEnv.step (multi_ue/multi_agent.py:149-179): only the current UE's entry of `actions` is
        applied (the agent dict holds that id only, multi_agent.py:21-30); rates and rewards are updated; after the last
        UE of the order the UEs move, the rates are updated again and time advances; then the NEXT UE becomes current and
        its observation row and its multi-agent reward are returned."""
        actions = np.asarray(actions)
        cur = self.ues[self.ue_order_idx]
        a = int(actions[self.ue_order_idx])
        if a > 0:
            self.connect_to_bs(cur, a - 1)
        rewards_before = self.update_ue_drs_rewards()
        moved = False
        if self.ue_order_idx + 1 < len(self.ues):
            self.ue_order_idx += 1
        else:
            self.ue_order_idx = 0
            self.last_lost_conn = [self.move(ue) for ue in self.ues]
            self.update_ue_drs_rewards(update_only=True)
            self.time += 1
            moved = True
        nxt = self.ue_order_idx
        o = self.get_ue_obs(self.ues[nxt])
        obs = np.concatenate([np.asarray(o[k], dtype=np.float64) for k in sorted(o.keys())])
        reward = self.step_reward(rewards_before)[nxt]
        out = self.snapshot()
        lost = self.last_lost_conn if moved else [0] * len(self.ues)
        out.update(obs=obs, reward=np.float64(reward), lost_conn=self._pad(np.asarray(lost, dtype=np.int32)),
                   sum_utility=np.float64(sum([self.utility(ue) for ue in self.ues])), time=self.time, done=None)
        return out

    def step(self, actions):
        """single_ue/base.py:413-466. actions: int[max_ues], 0 = noop, b+1 = toggle BS b (entry i = i-th UE present)"""
        actions = np.asarray(actions)
        assert actions.shape == (self.max_ues,) and np.all(actions >= 0) and np.all(actions <= self.n_bs)
        if self.sequential:
            return self.step_sequential(actions)
        for pos, ue in enumerate(self.ues):                     # base.py:247-282
            a = int(actions[pos])
            if a > 0:
                self.connect_to_bs(ue, a - 1)
        # base.py:429-443: arrivals / departures after the actions, before the rates and rewards
        if self.new_ue_interval is not None and self.time > 0 and self.time % self.new_ue_interval == 0:
            self.add_new_ue()
        if self.ue_arrival is not None and self.time in self.ue_arrival:
            n = self.ue_arrival[self.time]
            for _ in range(abs(n)):
                self.add_new_ue() if n > 0 else self.remove_ue()
            assert len(self.ues) <= self.max_ues
        rewards_before = self.update_ue_drs_rewards()
        self.last_lost_conn = [self.move(ue) for ue in self.ues]
        self.update_ue_drs_rewards(update_only=True)
        self.time += 1
        sum_utility = sum([self.utility(ue) for ue in self.ues])
        self.total_utility += sum_utility
        obs = self.get_obs()
        reward = self.step_reward(rewards_before)
        out = self.snapshot()
        out.update(obs=obs, reward=reward, lost_conn=self._pad(np.asarray(self.last_lost_conn, dtype=np.int32)),
                   sum_utility=np.float64(sum_utility), time=self.time, done=None)
        return out

    # ------------------------------------------------------------------ brute force (agent/brute_force.py)
    def test_ue_actions(self, actions):
        """single_ue/base.py:284-313: rewards of a joint action tried on the current state, then reverted"""
        original_ewma = [ue.ewma_dr for ue in self.ues]

        def apply():                                            # base.py:247-282
            for pos, ue in enumerate(self.ues):
                if int(actions[pos]) > 0:
                    self.connect_to_bs(ue, int(actions[pos]) - 1)
        apply()
        self.update_ue_drs_rewards(update_only=True)
        for ue in self.ues:                                     # user.py:148-157 update_ewma_dr
            ue.ewma_dr = 0.9 * self.curr_dr(ue) + (1 - 0.9) * ue.ewma_dr
        rewards = self.update_ue_drs_rewards()
        apply()                                                 # "to revert the action, apply it again"
        for ue, e in zip(self.ues, original_ewma):
            ue.ewma_dr = e
        self.update_ue_drs_rewards(update_only=True)
        return rewards

    def candidate_action(self, c):
        """brute_force.py:26-62: the c-th action = digits of c in base M + 1, most significant digit = first UE"""
        digits = []
        for _ in range(self.max_ues):
            digits.append(c % (self.n_bs + 1))
            c //= self.n_bs + 1
        return digits[::-1]

    def brute_force_rewards(self):
        """brute_force.py:64-94: central step reward of every joint action; the agent takes np.argmax (first maximum)"""
        assert self.kind == 'central'
        n_cand = (self.n_bs + 1) ** self.max_ues
        return np.array([float(self.step_reward(self.test_ue_actions(self.candidate_action(c)))) for c in range(n_cand)])

    # ------------------------------------------------------------------ snapshots (same keys as ref_loader.RefTrace)
    def snapshot(self):
        n, m = len(self.ues), self.n_bs
        mask = np.zeros((n, m), dtype=np.uint8)
        rates = np.zeros((n, m), dtype=np.float64)
        snr = np.zeros((n, m), dtype=np.float64)
        for j, ue in enumerate(self.ues):
            for b, r in ue.bs_dr.items():
                mask[j, b] = 1
                rates[j, b] = r
            for b in range(m):
                snr[j, b] = self.snr(b, ue)
        d = dict(
            pos=np.array([[ue.x, ue.y] for ue in self.ues], dtype=np.float64), mask=mask, link_rates=rates, snr=snr,
            curr_dr=np.array([float(self.curr_dr(ue)) for ue in self.ues]),
            ewma=np.array([float(ue.ewma_dr) for ue in self.ues]),
            utility=np.array([float(self.utility(ue)) for ue in self.ues]),
            # RandomWaypoint: velocity, waypoint, pausing, curr_pause; UniformMovement: move_x, move_y, -1, 0, 0
            movement=np.array([[ue.move_x, ue.move_y, -1.0, 0.0, 0.0] if getattr(ue, 'uniform', None) is not None else
                               [ue.velocity, ue.wx, ue.wy, float(ue.pausing), ue.curr_pause] for ue in self.ues],
                              dtype=np.float64))
        d = {k: self._pad(v) for k, v in d.items()}
        d['num_ue'] = n
        return d

    def reset_trace(self):
        obs = self.reset()
        out = self.snapshot()
        if self.sequential:                                     # multi_agent.py:122-124: the current UE's row only
            obs = obs[self.ue_order_idx]
        out['obs'] = obs
        return out
