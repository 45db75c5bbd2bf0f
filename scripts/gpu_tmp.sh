cd $GRAFT_REPO_ROOT
python - <<'PY'
import time, torch, numpy as np
from deepcomp_b200 import BatchedMobileEnv
bs, wh = [(10, 10), (110, 10), (60, 96.60254037844386)], (120, 106)
for n_ue in (10, 12):
    env = BatchedMobileEnv(num_envs=1, n_ue=n_ue, bs_xy=bs, map_wh=wh, kind='central', seeds=[5], reward='avg')
    env.reset()
    env.test_actions(0, 0, 1000); torch.cuda.synchronize()
    t0 = time.perf_counter(); r = env.test_actions(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'{n_ue} UEs x 3 BS: {env.num_joint_actions} joint actions in {dt*1e3:.2f} ms = {env.num_joint_actions/dt:.3e} candidates/s')
PY
