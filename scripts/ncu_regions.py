"""Instruction / stall-sample share per `// [region:NAME]` block of dcb_step_body.cuh from an .ncu-rep.

    python scripts/ncu_regions.py gpurun_out/prof.ncu-rep 'dcb_step_kernelILi768'
"""
import contextlib
import importlib.util
import io
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    spec = importlib.util.spec_from_file_location('h', os.path.join(HERE, 'ncu_hotlines.py'))
    h = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(h)
    sys.argv = ['x', rep, os.path.join(ROOT, 'deepcomp_b200', 'libdeepcomp_b200.so'), pat, '100000']
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        h.main()
    lines = buf.getvalue().splitlines()
    src = open(os.path.join(ROOT, 'deepcomp_b200', 'csrc', 'dcb_step_body.cuh')).read().splitlines()
    marks = [(i + 1, m.group(1)) for i, l in enumerate(src) for m in [re.search(r'\[region:([^\]]+)\]', l)] if m]
    agg = {}
    for ln in lines[1:]:
        m = re.match(r'\s*([\d.]+)% inst\s+([\d.]+)% samples\s+(\S+):(\d+)', ln)
        if not m:
            continue
        pi, ps, f, l = float(m.group(1)), float(m.group(2)), m.group(3), int(m.group(4))
        key = f
        if f == 'dcb_step_body.cuh':
            key = 'layout'
            for start, name in marks:
                if l >= start:
                    key = name
        a = agg.setdefault(key, [0.0, 0.0])
        a[0] += pi
        a[1] += ps
    print(lines[0])
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f'{v[0]:5.1f}% inst {v[1]:5.1f}% samples  {k}')


if __name__ == '__main__':
    main()
