"""Multi-GPU plumbing: one process per GPU, envs sharded in contiguous blocks, no
env-side communication.  The reference has no distributed path at all (its `--horovod` flag is parsed
at humanoid/utils/helpers.py:179-182 and never read); these are the three exchanges a data-parallel PPO
run over the sharded envs needs, all latency-bound and far below NVLink bandwidth:

  * PPO gradients  -> one flat-bucket all-reduce per optimiser step (`FlatGradAllReduce`)
  * advantage mean / unbiased std -> all-reduce of (count, sum, sum of squares) between the two GAE
    kernels (`ti5_isaacgym_b200.algo.rollout_storage.gae_returns_(group=...)`)
  * adaptive-LR KL statistic -> `all_reduce_mean`, so every rank takes the same schedule branch
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style initialisation (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*); returns (rank, world, local)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": torch.device(f"cuda:{local}")} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_envs(total_envs, rank, world):
    """Contiguous block [start, start + count) of the global env index space owned by `rank`."""
    base, rem = divmod(total_envs, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def all_reduce_mean(t, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
        t /= dist.get_world_size(group)
    return t


def merge_moments(stats, group=None):
    """All-reduce (count, sum, sum of squares); returns (mean, unbiased std) of the union — the
    quantities `RolloutStorage.compute_returns` normalises with (rollout_storage.py:119)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, group=group)
    n, s, q = (float(x) for x in stats[:3])
    mean = s / n
    var = max((q - s * mean) / (n - 1.0), 0.0)
    return mean, var ** 0.5


class FlatGradAllReduce:
    """Averages the gradients of `module` across ranks with ONE all-reduce of a flat bucket.  Call
    `reduce()` between `loss.backward()` and `clip_grad_norm_` (dh_ppo.py:180-181); the ~0.86 M
    parameters of ActorCriticDH are 3.4 MB, so a single latency-bound collective per step is the
    right shape (no bucketing / overlap machinery)."""

    def __init__(self, module, group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)

    def reduce(self):
        if not (dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return
        o = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[o:o + n].zero_()
            else:
                self.flat[o:o + n].copy_(p.grad.reshape(-1))
            o += n
        dist.all_reduce(self.flat, group=self.group)
        self.flat /= dist.get_world_size(self.group)
        o = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = self.flat[o:o + n].view_as(p).clone()
            else:
                p.grad.copy_(self.flat[o:o + n].view_as(p))
            o += n
