"""Build libdeepcomp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libdeepcomp_b200.so')
OBJ_DIR = os.path.join(HERE, 'build')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    # the reference's arithmetic has no fused multiply-adds except the one inside np.linalg.norm, which the
    # kernels spell out with fma(); contraction would move UE trajectories by an ulp and break mask parity
    '--fmad=false',
    '-Xcompiler', '-fPIC',
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def headers():
    return (glob.glob(os.path.join(CSRC, '*.h')) + glob.glob(os.path.join(CSRC, '*.cuh')) +
            [os.path.join(os.path.dirname(HERE), 'include', 'deepcomp_b200.h')])


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in sources() + headers())


def build(force=False, verbose=False, defines=(), out=None):
    """One nvcc -c per translation unit (in parallel; unchanged units are reused), then one link step."""
    out = out or LIB
    if not force and out == LIB and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    tag = ''.join(sorted(d.replace('=', '-') for d in defines))
    obj_dir = os.path.join(OBJ_DIR, tag or 'default')
    os.makedirs(obj_dir, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in headers())
    flags = NVCC_FLAGS + [f'-D{d}' for d in defines] + (['-Xptxas', '-v'] if verbose else [])

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + '.o')
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and os.path.getmtime(obj) > hdr_t):
            return obj
        subprocess.check_call([nvcc] + flags + ['-c', '-o', obj, src])
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    subprocess.check_call([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', out] + objs)
    return out


if __name__ == '__main__':
    defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith('-D'))
    outs = [a[2:] for a in sys.argv[1:] if a.startswith('-o')]
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, defines=defs, out=outs[0] if outs else None))
