"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: env sharding, the global advantage
statistics that ride between the two GAE kernels, and the flat-bucket gradient all-reduce."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ti5_isaacgym_b200.distributed import FlatGradAllReduce, all_reduce_mean, merge_moments, shard_envs


def test_shard_envs_partitions_exactly():
    for total, world in ((65536, 8), (8192, 1), (1000, 3), (7, 8)):
        blocks = [shard_envs(total, r, world) for r in range(world)]
        assert sum(c for _, c in blocks) == total
        assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    adv = torch.randn(24, 64, 1, generator=g, dtype=torch.float64)          # the global batch, same on every rank
    start, count = shard_envs(64, rank, world)
    mine = adv[:, start:start + count]
    stats = torch.tensor([mine.numel(), mine.sum(), (mine * mine).sum(), 0.0], dtype=torch.float64)
    mean, std = merge_moments(stats)
    ok_moments = abs(mean - adv.mean().item()) < 1e-12 and abs(std - adv.std().item()) < 1e-12
    # gradient averaging
    torch.manual_seed(0)
    net = torch.nn.Linear(5, 3)
    x = torch.full((2, 5), float(rank + 1))
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    FlatGradAllReduce(net).reduce()
    gathered = [torch.zeros_like(local[0]) for _ in range(world)]
    dist.all_gather(gathered, local[0])
    ok_grads = torch.allclose(net.weight.grad, torch.stack(gathered).mean(0))
    kl = all_reduce_mean(torch.tensor([float(rank)]))
    ok_kl = abs(kl.item() - (world - 1) / 2) < 1e-12
    out[rank] = bool(ok_moments and ok_grads and ok_kl)
    dist.destroy_process_group()


def test_two_rank_gloo_exchanges():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, 29541 + os.getpid() % 400, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
