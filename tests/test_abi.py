"""CPU: the C-ABI library builds, loads and exports every symbol include/deepcomp_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from deepcomp_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'deepcomp_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(dcb_[a-z_0-9]+)\s*\(', hdr)))


def test_library_is_built_and_loads():
    build.build()
    assert os.path.exists(_lib.lib_path())
    L = _lib.load()
    assert L.dcb_abi_version() == _lib.DCB_ABI_VERSION


def test_every_declared_symbol_is_exported():
    L = ctypes.CDLL(_lib.lib_path())
    decl = declared_symbols()
    assert len(decl) >= 15
    for sym in decl:
        assert hasattr(L, sym), f"{sym} declared in include/deepcomp_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == decl, "python binding list and header are out of sync"


def test_struct_layouts_match_header():
    """ctypes mirrors of the ABI structs: sizes as the C compiler lays them out (LP64)."""
    assert ctypes.sizeof(_lib.DcbConfig) == 14 * 4 + 5 * 8
    assert ctypes.sizeof(_lib.DcbOutputs) == 6 * 8 + 6 * 8 + 7 * 8
    assert ctypes.sizeof(_lib.DcbStateHost) == 5 * 8
    assert ctypes.sizeof(_lib.DcbPolicy) == 2 * 4 + 8 + 2 * 8 + 8 + 8
    assert ctypes.sizeof(_lib.DcbObsVariant) == 2 * 4 + 8 + 4 * 4


def test_create_rejects_bad_arguments_without_touching_the_gpu():
    L = _lib.load()
    h = ctypes.c_void_p()
    cfg = _lib.DcbConfig(abi_version=999)
    assert L.dcb_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b'abi_version' in L.dcb_last_error()
    cfg = _lib.DcbConfig(abi_version=_lib.DCB_ABI_VERSION, num_envs=1, n_ue=1, n_bs=65)
    assert L.dcb_create(ctypes.byref(cfg), ctypes.byref(h)) == -2          # DCB_ERR_UNSUPPORTED
    assert b'n_bs' in L.dcb_last_error()
    assert L.dcb_create(None, ctypes.byref(h)) == -1


def test_product_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from deepcomp_b200 import BatchedMobileEnv
    with pytest.raises(RuntimeError, match='no CPU'):
        BatchedMobileEnv(num_envs=1, n_ue=2, bs_xy=[(10, 10)], map_wh=(120, 120))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'deepcomp_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f
                assert 'dcb_oracle' not in txt, f
