"""Whole-step device time with the L2 flushed before every step vs left warm: the gap is what cold misses cost, i.e. the
upper bound of what any further prefetching can win."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import make_cfg
from ti5_isaacgym_b200.envs import T1DHStandEnv
from ti5_isaacgym_b200.sim.synthetic import SimParams, fill_synthetic_state, synthetic_actions
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cfg = make_cfg(N)
env = T1DHStandEnv(cfg, SimParams(dt=cfg.sim.dt), 1, 'cuda:0', True, use_cuda_graph=True, materialize_obs=False)
gen = torch.Generator(device='cuda').manual_seed(1)
fill_synthetic_state(env.gym.tensors, env.env_origins, gen)
env.reset()
env.episode_length_buf = torch.randint(1, 2000, (N,), generator=gen, device='cuda')
act = synthetic_actions(N, gen, 'cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for do_flush in (1, 0, 1, 0):
    ts = []
    for i in range(120):
        if do_flush: flush.fill_(i & 255)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step(act); b.record()
        ts.append((a, b))
    torch.cuda.synchronize()
    v = sorted(x.elapsed_time(y) for x, y in ts[20:])
    print("flush=%d  median %.2f us  mean %.2f us" % (do_flush, v[len(v) // 2] * 1e3, sum(v) / len(v) * 1e3))
