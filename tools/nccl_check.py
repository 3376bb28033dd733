"""torchrun --nproc-per-node N tools/nccl_check.py — the three NCCL exchanges of a data-parallel PPO run over the sharded
envs, checked against their single-process meaning and timed on the device (max over ranks):
  * FlatGradAllReduce on a network with ActorCriticDH's parameter count (~0.86 M) vs an all_gather mean
  * gae_returns_(group=...) on a sharded (T, N) batch vs ti5_gae on the whole batch
  * all_reduce_mean of the KL statistic
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from ti5_isaacgym_b200.algo.rollout_storage import gae_returns_
from ti5_isaacgym_b200.distributed import FlatGradAllReduce, all_reduce_mean, init_from_env, shard_envs

rank, world, local = init_from_env("nccl")
torch.cuda.set_device(local)
dev = f"cuda:{local}"
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(3102 // 8, 768), torch.nn.ELU(), torch.nn.Linear(768, 512), torch.nn.ELU(),
                          torch.nn.Linear(512, 256), torch.nn.ELU(), torch.nn.Linear(256, 12)).to(dev)     # ~0.83 M parameters
sync = FlatGradAllReduce(net)
x = torch.randn(256, 3102 // 8, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank))
ok = True
for mode in ("fresh grads", "bucket views kept"):
    times = []
    for it in range(8):
        if mode == "fresh grads":
            net.zero_grad(set_to_none=True)
        else:
            net.zero_grad(set_to_none=False)
        net(x).square().mean().backward()
        local_grads = [p.grad.clone() for p in net.parameters()]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        a.record(); sync.reduce(); b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t))
        for p, g in zip(net.parameters(), local_grads):
            gathered = [torch.empty_like(g) for _ in range(world)]
            dist.all_gather(gathered, g)
            ok &= torch.allclose(p.grad, torch.stack(gathered).mean(0), rtol=1e-5, atol=1e-7)
    if rank == 0:
        print(f"grad all-reduce ({sum(p.numel() for p in net.parameters())} params, {mode}): "
              f"{1e3 * sorted(times)[len(times) // 2]:.1f} us per reduce(), max over {world} ranks")
# GAE with global advantage statistics
T, N = 24, 8192 * world
g = torch.Generator().manual_seed(99)
rew, val = torch.randn(T, N, 1, generator=g), torch.randn(T, N, 1, generator=g)
done = (torch.rand(T, N, 1, generator=g) < 0.02).byte()
last = torch.randn(N, 1, generator=g)
s, c = shard_envs(N, rank, world)
mine = [t[:, s:s + c].contiguous().to(dev) for t in (rew, val, done)] + [last[s:s + c].contiguous().to(dev)]
ret, adv = torch.empty_like(mine[0]), torch.empty_like(mine[0])
gae_returns_(mine[0], mine[1], mine[2], mine[3], ret, adv, 0.994, 0.9, None, dist.group.WORLD)
whole = [t.to(dev) for t in (rew, val, done, last)]
ret1, adv1 = torch.empty_like(whole[0]), torch.empty_like(whole[0])
gae_returns_(whole[0], whole[1], whole[2], whole[3], ret1, adv1, 0.994, 0.9, None, None)
ok &= torch.equal(ret, ret1[:, s:s + c]) and torch.allclose(adv, adv1[:, s:s + c], rtol=1e-5, atol=1e-6)
kl = all_reduce_mean(torch.tensor([float(rank)], device=dev))
ok &= abs(float(kl) - (world - 1) / 2) < 1e-6
flag = torch.tensor([1.0 if ok else 0.0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("nccl exchanges", "OK" if float(flag) == 1.0 else "MISMATCH", f"(world {world})")
dist.destroy_process_group()
sys.exit(0 if float(flag) == 1.0 else 1)
