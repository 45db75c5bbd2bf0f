"""CPU, world_size 2 over gloo: the rollout hand-off collective and the sharding rule (no GPU, no kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepcomp_b200 import env_seeds
from deepcomp_b200.distributed import gather_rollout, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(total, world, rank)
        T, N = 3, 4
        # slab [T, K_local, N]: value encodes (t, global env, ue) so misplaced rows are detectable
        t = torch.arange(T).view(T, 1, 1) * 10000 + torch.arange(lo, hi).view(1, -1, 1) * 10 + torch.arange(N).view(1, 1, N)
        full = gather_rollout(t.float(), total_envs=total, env_dim=1)
        want = (torch.arange(T).view(T, 1, 1) * 10000 + torch.arange(total).view(1, -1, 1) * 10
                + torch.arange(N).view(1, 1, N)).float()
        ok = full.shape == want.shape and bool(torch.equal(full, want))
        seeds = env_seeds(7, hi - lo, N, first_env=lo)
        q.put((rank, ok, seeds.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('total', [8, 7])
def test_gather_rollout_world2(total):
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    # the union of the shards' seeds is the unsharded batch's seed vector
    assert res[0][2] + res[1][2] == env_seeds(7, total, 4).tolist()
