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
    right shape (no bucketing / overlap machinery).

    Every parameter owns a view of the bucket.  After `reduce()` the parameters' `.grad` ARE those views; if the
    optimiser keeps them (`zero_grad(set_to_none=False)`) the next backward accumulates straight into the bucket
    and `reduce()` is the collective alone.  With the default `zero_grad()` (grads dropped, dh_ppo.py:179) the fresh
    gradients are gathered into the bucket by one multi-tensor copy.  NCCL averages inside the collective."""

    def __init__(self, module, group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        self.views, o = [], 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    def reduce(self):
        if not (dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return
        src, dst = [], []
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr() or g.stride() != v.stride():
                src.append(g)
                dst.append(v)
        if dst:
            torch._foreach_copy_(dst, src)
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(self.flat, group=self.group)
            self.flat /= dist.get_world_size(self.group)
        for p, v in zip(self.params, self.views):
            p.grad = v
