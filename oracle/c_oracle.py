"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/dcb_oracle.c (the fast C restatement).

Never imported by the product path.  ``build()`` compiles the C file with gcc (no CUDA, no reference sources).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, 'dcb_oracle.c')
_LIB = os.path.join(_HERE, 'libdcb_oracle.so')

SHARING_CODE = {'resource-fair': 0, 'rate-fair': 1, 'proportional-fair': 2, 'max-cap': 3}
SHARING_MIX = ['resource-fair', 'rate-fair', 'proportional-fair']
REWARD_CODE = {'avg': 0, 'sum': 1, 'min': 2}
KIND_CODE = {'central': 0, 'multi': 1}

_lib = None


def build(force=False):
    """gcc -O2 -ffp-contract=off (no implicit FMA: the restatement places its one FMA explicitly) -fopenmp"""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        cmd = ['gcc', '-O2', '-ffp-contract=off', '-fopenmp', '-shared', '-fPIC', '-o', _LIB, _SRC, '-lm']
        subprocess.check_call(cmd)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, ip, dp, dp,
                                 ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_set_variants.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int]
        L.orc_set_interference.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_reset.argtypes = [ctypes.c_void_p]
        L.orc_step.argtypes = [ctypes.c_void_p, ip]
        L.orc_obs_size.argtypes = [ctypes.c_void_p]
        L.orc_reward_size.argtypes = [ctypes.c_void_p]
        L.orc_get.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 13
        L.orc_batch_run.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ip, ctypes.c_int, ctypes.c_int]
        L.orc_batch_trace.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ip, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_rng_draws.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ip, ip, ip]
        L.orc_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _iptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def vel_spec(velocities, n_ue):
    if not isinstance(velocities, (list, tuple)):
        velocities = [velocities] * n_ue
    return np.array([-1.0 if v == 'slow' else (-2.0 if v == 'fast' else float(v)) for v in velocities],
                    dtype=np.float64)


def sharing_codes(sharing, n_bs):
    if isinstance(sharing, str):
        sharing = [sharing if sharing != 'mixed' else SHARING_MIX[b % 3] for b in range(n_bs)]
    return np.array([SHARING_CODE[s] for s in sharing], dtype=np.int32)


def init_xy(init_pos, n_ue):
    out = np.full((n_ue, 2), np.nan, dtype=np.float64)
    if init_pos is not None:
        for i, (x, y) in enumerate(init_pos):
            if x != 'random':
                out[i, 0] = x
            if y != 'random':
                out[i, 1] = y
    return out


def rng_draws(seed, n_raw, ranges):
    """Known-answer helper: first n_raw raw 32-bit outputs and randint(lo,hi) draws of random.Random(seed)."""
    L = lib()
    raw = np.zeros(n_raw, dtype=np.uint32)
    lo = np.array([r[0] for r in ranges], dtype=np.int32)
    hi = np.array([r[1] for r in ranges], dtype=np.int32)
    out = np.zeros(len(ranges), dtype=np.int32)
    L.orc_rng_draws(seed, n_raw, raw.ctypes.data, len(ranges), _iptr(lo), _iptr(hi), _iptr(out))
    return raw, out


class COracleEnv:
    """Same constructor / trace keys as oracle.deepcomp_oracle.OracleEnv."""

    def __init__(self, kind, n_ue, bs_xy, map_wh, sharing='mixed', velocities='slow', seed=None, reward='avg',
                 episode_length=100, rand_episodes=False, init_pos=None, pause_duration=2, border_buffer=10,
                 util_func='log', dr_req=1, obs_norm='rel', interference=False):
        """`interference`: EXTENSION, not in the reference (SNR only) -- this restatement is its only oracle."""
        assert util_func in ('log', 'step') and obs_norm in ('rel', 'max')
        self.L = lib()
        self.kind, self.n_ue, self.n_bs = kind, n_ue, len(bs_xy)
        bs = np.ascontiguousarray(np.asarray(bs_xy, dtype=np.float64).reshape(-1, 2))
        sh = sharing_codes(sharing, self.n_bs)
        vs = vel_spec(velocities, n_ue)
        ixy = init_xy(init_pos, n_ue)
        self.h = ctypes.c_void_p(self.L.orc_create(
            KIND_CODE[kind], n_ue, self.n_bs, _dptr(bs), int(map_wh[0]), int(map_wh[1]), _iptr(sh), _dptr(vs),
            _dptr(ixy), 0 if seed is None else int(seed), int(seed is not None), REWARD_CODE[reward],
            int(rand_episodes), pause_duration, border_buffer))
        self.obs_size = self.L.orc_obs_size(self.h)
        self.reward_size = self.L.orc_reward_size(self.h)
        if util_func != 'log' or obs_norm != 'rel':
            self.L.orc_set_variants(self.h, int(util_func == 'step'), float(dr_req), int(obs_norm == 'max'))
        if interference:
            self.L.orc_set_interference(self.h, 1)

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def _trace(self, with_step):
        n, m = self.n_ue, self.n_bs
        pos = np.zeros((n, 2)); mask = np.zeros((n, m), dtype=np.uint8); rates = np.zeros((n, m))
        snr = np.zeros((n, m)); curr = np.zeros(n); ewma = np.zeros(n); util = np.zeros(n); mov = np.zeros((n, 5))
        obs = np.zeros(self.obs_size); rew = np.zeros(self.reward_size); lost = np.zeros(n, dtype=np.int32)
        su = ctypes.c_double(0); tm = ctypes.c_int(0)
        self.L.orc_get(self.h, pos.ctypes.data, mask.ctypes.data, rates.ctypes.data, snr.ctypes.data,
                       curr.ctypes.data, ewma.ctypes.data, util.ctypes.data, mov.ctypes.data, obs.ctypes.data,
                       rew.ctypes.data, lost.ctypes.data, ctypes.addressof(su), ctypes.addressof(tm))
        if self.kind == 'multi':
            obs = obs.reshape(n, 4 * m + 1)
        out = dict(pos=pos, mask=mask, link_rates=rates, snr=snr, curr_dr=curr, ewma=ewma, utility=util,
                   movement=mov, obs=obs)
        if with_step:
            out.update(reward=rew[0] if self.kind == 'central' else rew, lost_conn=lost, sum_utility=su.value,
                       time=tm.value, done=None)
        return out

    def reset_trace(self):
        self.L.orc_reset(self.h)
        return self._trace(False)

    def step(self, actions):
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.int32))
        assert a.shape == (self.n_ue,)
        self.L.orc_step(self.h, _iptr(a))
        return self._trace(True)


def batch_run(envs, actions, nthreads=0):
    """actions: int32 [T, K, N]; steps every env T times using OpenMP threads."""
    L = lib()
    a = np.ascontiguousarray(actions, dtype=np.int32)
    T, K, _ = a.shape
    assert K == len(envs)
    arr = (ctypes.c_void_p * K)(*[e.h for e in envs])
    L.orc_batch_run(arr, K, _iptr(a), T, nthreads)


def batch_trace(envs, actions, pos_every, nthreads=0):
    """As batch_run, returning per step the connection bitmasks uint64 [T, K, N], the lost-link counts uint8 [T, K, N],
    the rewards float64 [T, K, R], and the positions after every `pos_every`-th step float64 [T // pos_every, K, N, 2]."""
    L = lib()
    a = np.ascontiguousarray(actions, dtype=np.int32)
    T, K, N = a.shape
    assert K == len(envs)
    R = envs[0].reward_size
    mask = np.zeros((T, K, N), dtype=np.uint64)
    lost = np.zeros((T, K, N), dtype=np.uint8)
    rew = np.zeros((T, K, R), dtype=np.float64)
    pos = np.zeros((T // pos_every, K, N, 2), dtype=np.float64)
    arr = (ctypes.c_void_p * K)(*[e.h for e in envs])
    L.orc_batch_trace(arr, K, _iptr(a), T, nthreads, mask.ctypes.data, lost.ctypes.data, rew.ctypes.data,
                      pos.ctypes.data, pos_every)
    return mask, lost, rew, pos
