"""Stall-reason samples of the step kernel per warp role (physics / observer) from an .ncu-rep source page.

    python scripts/ncu_stalls.py gpurun_out/prof.ncu-rep deepcomp_b200/libdeepcomp_b200.so 'dcb_step_kernelILi704ELb1'
"""
import csv
import os
import re
import subprocess
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ncu_hotlines as H  # noqa: E402


def main():
    rep, lib, pat = sys.argv[1:4]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
    hdr = rows[h]
    cols = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
    body = [r for r in rows[h + 1:] if len(r) == len(hdr)]
    sl = H.sass_lines(lib, pat)
    src = open(os.path.join(os.path.dirname(HERE), 'deepcomp_b200', 'csrc', 'dcb_step_body.cuh')).read().splitlines()
    marks = [(i + 1, m.group(1)) for i, l in enumerate(src) for m in [re.search(r'\[region:([^\]]+)\]', l)] if m]

    def region(f, l):
        if f != 'dcb_step_body.cuh':
            return None
        key = None
        for start, name in marks:
            if l >= start:
                key = name
        return key
    regs = [region(f, l) for f, l, _ in sl]
    first_o = next(i for i, r in enumerate(regs) if r and r.startswith('O.'))
    tot = defaultdict(lambda: defaultdict(int))
    for idx, r in enumerate(body):
        role = 'physics' if idx < first_o else 'observer'
        for c in cols:
            tot[role][hdr[c]] += int(r[c] or 0)
    for role, d in tot.items():
        s = sum(d.values())
        print(f'== {role}: {s} samples')
        for k, v in sorted(d.items(), key=lambda kv: -kv[1]):
            if v:
                print(f'   {100 * v / s:5.1f}%  {k}')


if __name__ == '__main__':
    main()
