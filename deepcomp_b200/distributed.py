"""Multi-GPU plumbing: env batches shard one contiguous slab per rank; the only collective is the rollout hand-off.

Env instances are independent: rank r of G owns global envs [r*K/G, (r+1)*K/G) and seeds them
by GLOBAL env index, so the union of the shards is bit-identical to the unsharded batch.  Nothing crosses GPUs inside
``step``.  When the PPO learner wants the fragment, ``gather_rollout`` all-gathers the per-rank slabs over NCCL
(NVLink 5 / NVSwitch); on CPU test rigs the same code runs over gloo.
"""
import torch
import torch.distributed as dist


def shard_bounds(total_envs, world_size, rank):
    """Contiguous slab [lo, hi) of rank `rank`; the first total % world ranks get one extra env."""
    base, rem = divmod(int(total_envs), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def make_sharded_env(total_envs, rank=None, world_size=None, base_seed=0, **scenario):
    """BatchedMobileEnv over this rank's slab, on cuda:LOCAL device, seeded by global env index."""
    from .batched import BatchedMobileEnv
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(total_envs, world_size, rank)
    return BatchedMobileEnv(num_envs=hi - lo, seed=base_seed, first_env=lo, **scenario)


def gather_rollout(local, total_envs=None, env_dim=1, group=None):
    """
    All-gather a rollout slab whose env axis is `env_dim` (e.g. obs [T, K/G, N, 4M+1] -> [T, K, N, 4M+1]).
    Equal shards use one all_gather_into_tensor; ragged shards (K % G != 0) fall back to all_gather of padded slabs.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    x = local.movedim(env_dim, 0).contiguous()
    if total_envs is None or total_envs % world == 0:
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x, group=group)
    else:
        sizes = [shard_bounds(total_envs, world, r) for r in range(world)]
        kmax = max(hi - lo for lo, hi in sizes)
        pad = torch.zeros((kmax,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[: x.shape[0]] = x
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out = torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
    return out.movedim(0, env_dim)
