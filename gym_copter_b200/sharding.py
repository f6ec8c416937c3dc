"""
Multi-GPU plumbing.  The env batch shards by contiguous global env id, one shard per rank
(= per GPU), and NOTHING is exchanged per step: every kernel input of env i is keyed by its
global id (Philox counter), so the union of the shards is bit-identical to the unsharded
batch.  The only collective is the optional SUM all-reduce of the episode-statistics vector.
"""

import os


def shard_range(global_envs, rank, world):
    """Contiguous [lo, hi) of global env ids owned by `rank`; sizes differ by at most 1 and
    every boundary is a multiple of 256 when global_envs >= 256 * world (keeps each shard's
    sub-buffers 16-byte aligned if the caller slices one big allocation)."""
    if not (0 <= rank < world) or global_envs < 0:
        raise ValueError('bad shard request rank=%r world=%r envs=%r' % (rank, world, global_envs))
    if global_envs >= 256 * world:
        tiles = (global_envs + 255) // 256
        lo_t = tiles * rank // world
        hi_t = tiles * (rank + 1) // world
        return min(lo_t * 256, global_envs), min(hi_t * 256, global_envs)
    return global_envs * rank // world, global_envs * (rank + 1) // world


def env_from_torchrun(default_world=1):
    """(rank, local_rank, world) from the torchrun environment."""
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')),
            int(os.environ.get('WORLD_SIZE', str(default_world))))


def all_reduce_stats(vec, group=None):
    """SUM all-reduce of a statistics tensor over the ranks (NCCL on GPUs, gloo in CPU tests).
    64-128 bytes, latency-bound, off the step path."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec


def make_sharded_env(variant, global_envs, rank, world, **kw):
    """The shard of a `global_envs`-wide batch that lives on this rank's GPU."""
    from .envs import CopterVecEnv
    lo, hi = shard_range(global_envs, rank, world)
    return CopterVecEnv(variant, hi - lo, env_offset=lo, **kw)
