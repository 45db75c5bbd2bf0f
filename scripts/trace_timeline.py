"""Phase timeline of CTA 0 of the fused step kernel from a -DDCB_TRACE build (clock64 at phase boundaries).

    DCB_LIB_PATH=$PWD/gpurun_exp_TRACE.so python scripts/trace_timeline.py
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import grid_layout  # noqa: E402
from deepcomp_b200 import BatchedMobileEnv, _lib  # noqa: E402

K, N, M = 1024, 50, 10
W, H, bs = grid_layout(M)
env = BatchedMobileEnv(num_envs=K, n_ue=N, bs_xy=bs, map_wh=(W, H), kind='multi', seed=1000, episode_length=100)
a = torch.randint(0, M + 1, (100, K, N), dtype=torch.int32, device='cuda', generator=torch.Generator('cuda').manual_seed(0))
env.reset()
out = env.step_many(a)
env.reset()
env.step_many(a, out=out)
torch.cuda.synchronize()
L = _lib.load()
buf = np.zeros(2 * 16 * 8 * 8, dtype=np.int64)
L.dcb_trace_read.argtypes = [ctypes.c_void_p]
assert L.dcb_trace_read(buf.ctypes.data) == 0
tr = buf.reshape(2, 16, 8, 8)
nw = env.launch_geometry['threads'] // 64
names = {0: ['top', 'pre-move', 'pre-links', 'post-links', 'bar_or', '', 'bar2', 'end'],
         1: ['full', 'util-red', 'dense', 'staging', 'reward', 'end', '', '']}
t0 = tr[0, :nw, 0, 0].min()
for role, pts in ((0, [0, 1, 2, 3, 4, 6, 7]), (1, [0, 1, 2, 3, 4, 5])):
    print('== physics' if role == 0 else '== observer', '(cycles since the first physics warp entered step 40; rows: warps)')
    for step in range(3):
        print(f' step {40 + step}: ' + ' '.join(f'{names[role][p]:>10s}' for p in pts))
        for w in range(nw):
            print('          ' + ' '.join(f'{tr[role, w, step, p if p != 6 else 5] - t0:10d}' for p in pts))
per_step = (tr[0, :nw, 7, 0] - tr[0, :nw, 0, 0]) / 7.0
print('cycles per step (physics top to top):', per_step.mean())

# ---- per-CTA wall time of the traced launch (globaltimer): the kernel ends with its slowest CTA
cta = np.zeros(3 * 4096, dtype=np.int64)
L.dcb_trace_read_cta.argtypes = [ctypes.c_void_p]
assert L.dcb_trace_read_cta(cta.ctypes.data) == 0
g = env.launch_geometry['grid']
t0 = cta[0:3 * g:3]
t1 = cta[1:3 * g:3]
sm = cta[2:3 * g:3]
dur = (t1 - t0) / 1e3
order = np.argsort(dur)
print(f'per-CTA duration of one 100-step launch (us): min {dur.min():.1f} median {np.median(dur):.1f} '
      f'p90 {np.percentile(dur, 90):.1f} max {dur.max():.1f}; launch span {(t1.max() - t0.min()) / 1e3:.1f}')
print('slowest CTAs (cta, sm, us):', [(int(c), int(sm[c]), round(float(dur[c]), 1)) for c in order[-8:]])
print('fastest CTAs (cta, sm, us):', [(int(c), int(sm[c]), round(float(dur[c]), 1)) for c in order[:8]])
