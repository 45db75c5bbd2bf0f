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


def _sass_of(cubin_name, tmp):
    """SASS of one cubin of the built library (cuobjdump -xelf / -sass), split by function"""
    import shutil
    import subprocess
    if not shutil.which('cuobjdump'):
        pytest.skip('cuobjdump not found')
    subprocess.check_call(['cuobjdump', '-xelf', cubin_name, os.path.abspath(_lib.lib_path())], cwd=tmp,
                          stdout=subprocess.DEVNULL)
    txt = subprocess.run(['cuobjdump', '-sass', os.path.join(tmp, cubin_name)], capture_output=True, text=True).stdout
    funcs, name = {}, None
    for ln in txt.splitlines():
        m = re.match(r'\s*Function : (\S+)', ln)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
            funcs[name].append(ln.split('*/', 1)[1].strip())
    assert funcs, f'no functions in {cubin_name}'
    return funcs


def _count(lines, mnemonic):
    return sum(1 for ln in lines if re.search(r'(^|\s)' + mnemonic + r'\b', ln))


def test_built_kernels_are_sm100a_and_use_the_instructions_the_design_names(tmp_path):
    """Static evidence from the shipped binary (no GPU): every cubin targets sm_100a, the fused step kernel's measured
    instance sends its observation tile with the TMA bulk store (UBLKCP.G.S), prefetches the waypoint table with LDGSTS,
    synchronises its two warp groups with named barriers and computes in fp64; nothing uses TMA tensor loads (the pair tile
    never exists in HBM); the wide kernel reduces with redux.sync / shuffles and its interference pass is a function of
    its own.  (profiles/r02_sass_tma_excerpt.txt is the human-readable form.)"""
    import subprocess
    elfs = subprocess.run(['cuobjdump', '-lelf', os.path.abspath(_lib.lib_path())], capture_output=True, text=True).stdout
    names = re.findall(r'ELF file\s+\d+:\s+(\S+)', elfs)
    assert names and all(n.endswith('.sm_100a.cubin') for n in names), names
    fused = _sass_of('dcb_step_k704.sm_100a.cubin', str(tmp_path))
    head = [f for f in fused if 'dcb_step_kernel_704ILb1ELb0ELb0' in f]          # <M32, !PAD, !CENTRAL>: the headline instance
    assert len(head) == 1, list(fused)
    k = fused[head[0]]
    assert _count(k, r'UBLKCP\.G\.S') >= 1 and _count(k, r'LDGSTS\S*') >= 1
    assert _count(k, r'BAR\.SYNC\S*') + _count(k, r'BAR\.ARV\S*') + _count(k, r'BAR\.RED\S*') >= 10
    assert _count(k, r'DFMA') > 200 and _count(k, r'MUFU\.\S+') > 10
    assert _count(k, r'UTMALDG\S*') == 0 and not any('wgmma' in ln.lower() or 'HMMA' in ln for ln in k)
    assert _count(k, r'STL\S*') + _count(k, r'LDL\S*') <= 24, 'the measured instance spills more than a few words'
    wide = _sass_of('dcb_wide.sm_100a.cubin', str(tmp_path))
    plain = [f for f in wide if 'dcb_wide_kernelILb0ELb0ELb0' in f]
    assert len(plain) == 1
    w = wide[plain[0]]
    assert _count(w, r'CREDUX\S*') + _count(w, r'REDUX\S*') >= 1 and _count(w, r'SHFL\S*') >= 10 and _count(w, r'VOTE\S*') >= 2
    # the interference pass is a function of its own in each of the four general (EXT) instances, and in no other
    syms = subprocess.run(['cuobjdump', '-symbols', os.path.join(str(tmp_path), 'dcb_wide.sm_100a.cubin')],
                          capture_output=True, text=True).stdout
    owners = re.findall(r'dcb_wide_kernelILb([01])ELb([01])ELb([01])EEEv8StepArgs\$\S*wide_interference_pass', syms)
    assert sorted(owners) == sorted((p, '1', c) for p in '01' for c in '01'), owners
