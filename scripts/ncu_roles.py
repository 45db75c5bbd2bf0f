"""Split the executed warp-instructions of the step kernel by warp role (physics / observer) and report, per role,
instructions per warp and step, the heaviest source lines and the stall-sample share.

    python scripts/ncu_roles.py gpurun_out/prof.ncu-rep deepcomp_b200/libdeepcomp_b200.so 'dcb_step_kernelILi704ELb1' [T] [warps_per_group] [ctas]

Roles are told apart by SASS address: everything after the first instruction attributed to the `[region:O.setup]`
block of dcb_step_body.cuh belongs to the observers, the shared prologue is reported on its own.
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
    T = int(sys.argv[4]) if len(sys.argv) > 4 else 100
    wpg = int(sys.argv[5]) if len(sys.argv) > 5 else 11
    ctas = int(sys.argv[6]) if len(sys.argv) > 6 else 147
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
    hdr = rows[h]
    ii, si = hdr.index('Instructions Executed'), hdr.index('# Samples')
    inst = [(int(r[ii]), int(r[si]), r[1]) for r in rows[h + 1:] if len(r) == len(hdr)]
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
    first_p = next(i for i, r in enumerate(regs) if r and r.startswith('P.'))
    first_o = next(i for i, r in enumerate(regs) if r and r.startswith('O.'))
    roles = ['prologue' if i < min(first_p, first_o) else ('physics' if i < first_o else 'observer') for i in range(len(sl))]
    tot = defaultdict(lambda: [0, 0])
    lines = defaultdict(lambda: defaultdict(lambda: [0, 0]))
    for (n, s, _), (f, l, _), role in zip(inst, sl, roles):
        tot[role][0] += n
        tot[role][1] += s
        lines[role][(f, l)][0] += n
        lines[role][(f, l)][1] += s
    all_i = sum(v[0] for v in tot.values())
    all_s = sum(v[1] for v in tot.values())
    for role in ('prologue', 'physics', 'observer'):
        n, s = tot[role]
        print(f'== {role}: {100 * n / all_i:.1f}% inst, {100 * s / all_s:.1f}% samples, '
              f'{n / (ctas * wpg * T):.0f} warp-instr per warp and step')
        for (f, l), (ni, sm) in sorted(lines[role].items(), key=lambda kv: -kv[1][0])[:int(os.environ.get('TOP', 25))]:
            text = ''
            if f == 'dcb_step_body.cuh' and l <= len(src):
                text = src[l - 1].strip()[:90]
            print(f'   {ni / (ctas * wpg * T):6.1f}/ws {100 * sm / all_s:5.1f}% smp  {f}:{l}  {text}')


if __name__ == '__main__':
    main()
