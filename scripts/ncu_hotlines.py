"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (needs -lineinfo at build time).

    python scripts/ncu_hotlines.py gpurun_out/prof.ncu-rep deepcomp_b200/libdeepcomp_b200.so 'dcb_step_kernelILi256' [top]

ncu's CSV source page lists SASS in address order; nvdisasm -g lists the same SASS with `//## File ..., line N`
markers.  The two are joined by instruction index within the kernel.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(lib, kernel_pat):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith('.cubin'):
            continue
        txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, cur_line, cur_file, in_fn = None, None, None, False
        for ln in txt.splitlines():
            m = re.match(r'\s*\.text\.(\S+):', ln)
            if m:
                cur_fn = m.group(1)
                in_fn = re.search(kernel_pat, cur_fn) is not None
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_file, cur_line = os.path.basename(m.group(1)), int(m.group(2))
                continue
            if in_fn and re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
                out.append((cur_file, cur_line, ln.split('*/', 1)[1].strip()))
        if out:
            break
    return out


def main():
    rep, lib, pat = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
    hdr = rows[h]
    ii, si = hdr.index('Instructions Executed'), hdr.index('# Samples')
    inst = [(int(r[ii]), int(r[si]), r[1]) for r in rows[h + 1:] if len(r) == len(hdr)]
    sl = sass_lines(lib, pat)
    if len(sl) != len(inst):
        print(f'warning: {len(sl)} SASS lines from nvdisasm vs {len(inst)} in the report', file=sys.stderr)
    agg = defaultdict(lambda: [0, 0])
    for (n, s, _), (f, l, _) in zip(inst, sl):
        agg[(f, l)][0] += n
        agg[(f, l)][1] += s
    tot_i = sum(v[0] for v in agg.values())
    tot_s = sum(v[1] for v in agg.values())
    src = {}
    print(f'total warp instructions {tot_i}, samples {tot_s}')
    for (f, l), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in src:
            for root in ('deepcomp_b200/csrc', '.'):
                pth = os.path.join(root, f or '')
                if os.path.exists(pth):
                    src[f] = open(pth).read().splitlines()
                    break
            else:
                src[f] = []
        text = src[f][l - 1].strip() if l and l <= len(src[f]) else ''
        print(f'{100 * n / tot_i:5.1f}% inst {100 * s / max(tot_s, 1):5.1f}% samples  {f}:{l}  {text[:100]}')


if __name__ == '__main__':
    main()
