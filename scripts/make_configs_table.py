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
                    ('cfg4c', 'config 4, central observation'), ('k256', 'sweep: 50 x 10 x 256'),
                    ('k4096', 'sweep: 50 x 10 x 4096'), ('k16384', 'sweep: 50 x 10 x 16384'),
                    ('k65536', 'sweep: 50 x 10 x 65536')]:
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


def dram(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, v = rr[0], rr[1], rr[2]

    def val(k):
        return float(v[hdr.index(k)]) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[hdr.index(k)]]
    return val('dram__bytes_read.sum'), val('dram__bytes_write.sum')


r, w = dram(f'{G}/prof_{T}.ncu-rep')
rw, ww = dram(f'{G}/prof_wide_{T}.ncu-rep')
json.dump({
    "multi:50x10x1024:F100": {
        "dram_bytes_per_launch": int(r + w), "algorithmic_bytes_per_launch": 11800 * 1024 * 100,
        "kernel": "dcb_step_kernel_704<true, false>, one 100-step fragment launch of the 50 UE x 10 BS x 1024 env batch",
        "source": f"profiles/{T}_step_kernel_ncu_summary.csv: ncu --set full, dram__bytes_read.sum ({r / 1e6:.2f} MB) + "
                  f"dram__bytes_write.sum ({w / 1e6:.2f} MB)"},
    "multi:1000x50x1024:F4": {
        "dram_bytes_per_launch": int(rw + ww), "algorithmic_bytes_per_launch": 884000 * 1024 * 4,
        "kernel": "dcb_wide_kernel<false>, one 4-step fragment launch of the 1000 UE x 50 BS x 1024 env batch",
        "source": f"profiles/{T}_wide_kernel_ncu_summary.csv: ncu --set full, dram__bytes_read.sum ({rw / 1e6:.2f} MB) + "
                  f"dram__bytes_write.sum ({ww / 1e6:.2f} MB)"},
}, open(f'{P}/traffic.json', 'w'), indent=1)
print(open(f'{P}/{T}_configs.md').read())
