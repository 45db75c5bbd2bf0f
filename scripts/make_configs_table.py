"""profiles/<tag>_configs.md + profiles/<tag>_bench_<cfg>.json + profiles/traffic.json from gpurun_out/ (scripts/gpu_final.sh)."""
import csv
import json
import subprocess
import sys

T = sys.argv[1]
G, P = 'gpurun_out', 'profiles'
rows = []
for name, label in [('cfg3', 'BASELINE config 3 per GPU: 200 UE x 20 BS x 512 envs, multi'),
                    ('cfg4', 'BASELINE config 4 per GPU: 1000 UE x 50 BS x 1024 envs, multi'),
                    ('cfg4c', 'config 4, central observation'),
                    ('cfg4i', 'config 4 with the interference extension (SINR; not in the reference, not parity-graded)'),
                    ('cfg3c', 'config 3, central observation'),
                    ('central', 'headline shape, CentralRelNormEnv: 50 UE x 10 BS x 1024 envs')]:
    d = json.load(open(f'{G}/bench_{T}_{name}.json'))
    json.dump(d, open(f'{P}/{T}_bench_{name}.json', 'w'))
    rows.append((label, d))
rows.insert(0, ('HEADLINE (BASELINE config 1 per GPU): 50 UE x 10 BS x 1024 envs, multi', json.load(open(f'{G}/bench_{T}.json'))))
with open(f'{P}/{T}_configs.md', 'w') as f:
    f.write(f'# {T}: bench lines of every BASELINE.json config on 1 x B200 (`scripts/gpu_configs.sh`, `bench.py`)\n\n')
    f.write('| workload | kernel | geometry (envs/CTA, threads, smem, grid) | env-steps/s | us / batched step | B / env-step | '
            'achieved GB/s | frac of measured 6554.2 GB/s | e2e env-steps/s |\n|---|---|---|---|---|---|---|---|---|\n')
    for label, d in rows:
        g, r = d['run']['launch_geometry'], d['roofline']
        f.write(f"| {label} | {r['kernel']} | {g['envs_per_cta']}, {g['threads']}, {g['smem_bytes']}, {g['grid']} | "
                f"{d['value']:.3e} | {1e3 * d['ms_per_step']:.2f} | {r['algorithmic_bytes_per_env_step']} | "
                f"{r['achieved']:.0f} | {100 * r['frac']:.1f} % | {d['e2e']['value']:.3e} |\n")


def dram(rep, dur_us=None):
    """DRAM bytes (read, write) of one captured launch: the first one, or the first whose duration lies in dur_us"""
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1, 'ms': 1e3, 'ns': 1e-3}
    for v in rr[2:]:
        def val(k):
            return float(v[hdr.index(k)]) * scale[units[hdr.index(k)]]
        if dur_us is None or dur_us[0] <= val('gpu__time_duration.sum') <= dur_us[1]:
            return val('dram__bytes_read.sum'), val('dram__bytes_write.sum')
    raise SystemExit(f'no launch of the wanted duration in {rep}')


r, w = dram(f'{G}/prof_{T}.ncu-rep')
r20, w20 = dram(f'{G}/prof_f20_{T}.ncu-rep', dur_us=(95.0, 170.0))      # the 20-step launches among the captured ones
rw, ww = dram(f'{G}/prof_wide_{T}.ncu-rep')
json.dump({
    "multi:50x10x1024:F20": {
        "dram_bytes_per_launch": int(r20 + w20), "algorithmic_bytes_per_launch": 11800 * 1024 * 20,
        "kernel": "dcb_step_kernel_704<true, false, false>, one 20-step launch (the fragment of `bench.py --steps 20`)",
        "source": f"profiles/{T}_step_kernel_f20_ncu_summary.csv: ncu --set full, dram__bytes_read.sum ({r20 / 1e6:.2f} MB) + "
                  f"dram__bytes_write.sum ({w20 / 1e6:.2f} MB)"},
    "multi:50x10x1024:F100": {
        "dram_bytes_per_launch": int(r + w), "algorithmic_bytes_per_launch": 11800 * 1024 * 100,
        "kernel": "dcb_step_kernel_704<true, false, false>, one 100-step fragment launch of the 50 UE x 10 BS x 1024 env batch",
        "source": f"profiles/{T}_step_kernel_ncu_summary.csv: ncu --set full, dram__bytes_read.sum ({r / 1e6:.2f} MB) + "
                  f"dram__bytes_write.sum ({w / 1e6:.2f} MB)"},
    "multi:1000x50x1024:F4": {
        "dram_bytes_per_launch": int(rw + ww), "algorithmic_bytes_per_launch": 884000 * 1024 * 4,
        "kernel": "dcb_wide_kernel<false>, one 4-step fragment launch of the 1000 UE x 50 BS x 1024 env batch",
        "source": f"profiles/{T}_wide_kernel_ncu_summary.csv: ncu --set full, dram__bytes_read.sum ({rw / 1e6:.2f} MB) + "
                  f"dram__bytes_write.sum ({ww / 1e6:.2f} MB)"},
}, open(f'{P}/traffic.json', 'w'), indent=1)
print(open(f'{P}/{T}_configs.md').read())

# env-batch sweep (BASELINE.json configs[4]) at 1 / 2 / 4 / 8 GPUs: gpurun_out/sweep_<tag>_n<N>.jsonl (scripts/gpu_sweep_multi.sh)
import os
sw = {}
for n in (1, 2, 4, 8):
    fn = f'{G}/sweep_{T}_n{n}.jsonl'
    if os.path.exists(fn):
        for ln in open(fn):
            if ln.startswith('{'):
                d = json.loads(ln)
                sw[(d['config']['envs_per_gpu'] * d['n_gpus'], n)] = d
if sw:
    totals = sorted({k[0] for k in sw})
    with open(f'{P}/{T}_sweep.md', 'w') as f:
        f.write(f'# {T}: env-batch sweep at 50 UE x 10 BS (BASELINE.json configs[4]): TOTAL envs split evenly over N GPUs\n\n'
                'Per cell: env-steps/s (whole job, device-timed, best of 3 repetitions) / roofline fraction per GPU / e2e '
                'env-steps/s through `dcb_step_many_host`. The 8192 row is the strong-scaling line of the north-star batch; the '
                'diagonal 1024 x N is the weak-scaling line. `scripts/gpu_sweep_multi.sh`.\n\n')
        f.write('| total envs | ' + ' | '.join(f'{n} GPU' + ('s' if n > 1 else '') for n in (1, 2, 4, 8)) + ' |\n|---|---|---|---|---|\n')
        for t in totals:
            cells = []
            for n in (1, 2, 4, 8):
                d = sw.get((t, n))
                cells.append('' if d is None else f"{d['value']:.3e} / {100 * d['roofline']['frac']:.1f} % / {d['e2e']['value']:.2e}")
            f.write(f'| {t} | ' + ' | '.join(cells) + ' |\n')
        f.write('\nDriver invocation (`bench.py --gpus N --steps 20 --warmup 5`, weak scaling, 1024 envs per GPU):\n\n'
                '| N | env-steps/s | roofline frac | avg launch ms | e2e env-steps/s | PCIe ceiling env-steps/s | NCCL nranks seen |\n|---|---|---|---|---|---|---|\n')
        for n in (1, 2, 4, 8):
            fn = f'{G}/bench_{T}_driver_n{n}.json'
            if not os.path.exists(fn):
                continue
            lines = [ln for ln in open(fn) if ln.startswith('{')]
            if not lines:
                continue
            d = json.loads(lines[-1])
            json.dump(d, open(f'{P}/{T}_bench_driver_flags_{n}gpu.json', 'w'))
            err = open(f'{G}/bench_{T}_driver_n{n}.err').read() if os.path.exists(f'{G}/bench_{T}_driver_n{n}.err') else ''
            seen = sorted(set(__import__('re').findall(r'nranks (\d+)', err)))
            f.write(f"| {n} | {d['value']:.4e} | {d['roofline']['frac']:.3f} | {d['roofline']['avg_launch_ms']:.4f} | "
                    f"{d['e2e']['value']:.3e} | {d['e2e']['pcie_ceiling']['value']:.3e} | {', '.join(seen) or '-'} |\n")
    print(open(f'{P}/{T}_sweep.md').read())
