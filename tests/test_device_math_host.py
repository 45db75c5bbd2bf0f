"""CPU: the table-driven fp64 math of the step kernels (deepcomp_b200/csrc/dcb_math.cuh), compiled for the host from the
very header the kernels include, against numpy / exact references.  Pins the accuracy DESIGN.md states for it -- <= 1e-13
relative to the reference's libm chain (station.py:110-127), i.e. four orders below the 1e-9 parity bar -- without a GPU."""
import ctypes
import os
import subprocess
from fractions import Fraction

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# station.py:26-30, 110-114 exactly as the reference (and dcb_create) evaluate them
CH = 0.8 + (1.1 * np.log10(2500.0) - 0.7) * 1.5 - 1.56 * np.log10(2500.0)
C1 = 69.55 + 26.16 * np.log10(2500.0) - 13.82 * np.log10(50.0) - CH
C2 = 44.9 - 6.55 * np.log10(50.0)
H = C2 / 20.0
C0 = np.log2(10.0) * (30.0 - C1) / 10.0 - np.log2(1e-9)


@pytest.fixture(scope='module')
def mh(tmp_path_factory):
    cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')
    if not os.path.exists(os.path.join(cuda_inc, 'cuda_runtime.h')):
        pytest.skip('CUDA headers not found')
    out = str(tmp_path_factory.mktemp('mh') / 'libdcb_math_host.so')
    subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-I', cuda_inc,
                           '-I', os.path.join(ROOT, 'deepcomp_b200', 'csrc'), '-x', 'c++',
                           os.path.join(ROOT, 'tests', 'native', 'dcb_math_host.cpp'), '-o', out])
    return ctypes.CDLL(out)


def tables():
    """dcb_api.cu: host_math_tables (inv, l2c, ex2, pwm, pwe) and DevParams::pw"""
    j = np.arange(16, dtype=np.float64)
    inv = 1.0 / (1.0 + (j + 0.5) / 16.0)
    t = np.concatenate([inv, -np.log2(inv), np.exp2(j / 16.0), inv ** H, np.exp2(C0 - H * j)])
    pw = np.ones(10)
    for k in range(1, 10):
        pw[k] = pw[k - 1] * (-H - (k - 1)) / k
    return np.ascontiguousarray(t), np.ascontiguousarray(pw)


def _call(fn, *arrays_and_scalars):
    args = []
    for a in arrays_and_scalars:
        args.append(a.ctypes.data_as(ctypes.c_void_p) if isinstance(a, np.ndarray) else a)
    fn(*args)


def _vec(fn, t, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    _call(fn, t, x, out, ctypes.c_int(x.size))
    return out


def test_log2_exp2_log1p_rcp(mh):
    t, _ = tables()
    rng = np.random.default_rng(0)
    x = np.exp(rng.uniform(np.log(1e-12), np.log(1e12), 200000))
    np.testing.assert_allclose(_vec(mh.mh_log2, t, x), np.log2(x), rtol=0, atol=2e-14)
    # near 1 the table form keeps its ABSOLUTE accuracy (l2c[j] + log2(1 + r) cancel), not a relative one: log_utility
    # (10 log10(dr) near dr = 1, utility.py:36-54) is compared with an absolute floor for that reason
    x1 = 1.0 + rng.uniform(-1e-3, 1e-3, 1000)
    np.testing.assert_allclose(_vec(mh.mh_log2, t, x1), np.log2(x1), rtol=0, atol=5e-15)
    y = rng.uniform(-300.0, 300.0, 200000)
    np.testing.assert_allclose(_vec(mh.mh_exp2, t, y), np.exp2(y), rtol=5e-15, atol=0)
    s = np.concatenate([np.exp(rng.uniform(np.log(1e-12), np.log(1e3), 100000)), [0.0, 0.03125, 2e-8, 7e-6]])
    # the reference rounds 1 + snr before the log (station.py:137): log2(fl(1 + s)), not log1p
    np.testing.assert_allclose(_vec(mh.mh_log2_1p, t, s), np.log2(1.0 + s), rtol=2e-13, atol=0)
    xr = np.exp(rng.uniform(np.log(1e-300), np.log(1e300), 100000))
    out = np.empty_like(xr)
    _call(mh.mh_rcp, xr, out, ctypes.c_int(xr.size))
    assert np.max(np.abs(out * xr - 1.0)) <= 4.5e-16                # <= 1 ulp after two Newton steps from a 20-bit seed


def test_snr_table_form_against_the_reference_chain_and_the_exact_power_law(mh):
    t, pw = tables()
    rng = np.random.default_rng(1)
    # every in-range link (d^2 <= 4750.5), the rest of the first table range, and the far pairs of the interference pass
    d2 = np.concatenate([rng.uniform(1.0, 4751.0, 150000), np.exp(rng.uniform(0.0, np.log(65536.0), 100000)),
                         np.exp(rng.uniform(np.log(65536.0), np.log(2.0 ** 29), 100000)),
                         [1.0, 68.92488308058006 ** 2, 65535.999999, 65536.0, 2.0 ** 20, 2.0 ** 29]])
    out = np.empty_like(d2)
    _call(mh.mh_snr, t, pw, ctypes.c_double(H), d2, out, ctypes.c_int(d2.size))
    # the reference's own chain (station.py:110-127) in numpy float64
    d = np.sqrt(d2)
    ref = 10.0 ** ((30.0 - (C1 + C2 * np.log10(d + 1e-16))) / 10.0) / 1e-9
    assert np.max(np.abs(out / ref - 1.0)) < 5e-14
    # ... and the exact power law 2^c0 * d2^-h in extended precision on a subsample
    sub = np.concatenate([np.arange(0, d2.size, 997), np.arange(d2.size - 6, d2.size)])
    ld = np.longdouble
    exact = np.exp2(ld(C0) - ld(H) * np.log2(d2[sub].astype(ld)))
    assert float(np.max(np.abs(out[sub].astype(ld) / exact - 1))) < 2e-14
    # the connection threshold distance of SURVEY 8c sits between two representable distances; the table form orders them
    lo, hi = 68.92488308058006, 68.92488308058007
    pair = np.array([lo * lo, hi * hi])
    o2 = np.empty(2)
    _call(mh.mh_snr, t, pw, ctypes.c_double(H), pair, o2, ctypes.c_int(2))
    assert o2[0] > o2[1] and abs(o2[0] / 2e-8 - 1.0) < 1e-12


def test_binomial_coefficients_are_the_series_of_the_power_law():
    """DevParams::pw = binom(-h, k): exact rational recurrence against the float one"""
    _, pw = tables()
    h = Fraction(H)
    c = Fraction(1)
    for k in range(1, 10):
        c = c * (-h - (k - 1)) / k
        assert abs(float(c) / pw[k] - 1.0) < 1e-14
    # truncation after degree 9 at |r| <= 1/32: ~5e-15 ("SNR / rate / utility / reward")
    c10 = c * (-h - 9) / 10
    assert abs(float(c10)) * (1.0 / 32.0) ** 10 < 5.5e-15
