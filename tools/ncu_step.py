"""A few production steps (Philox, direct launches, L2 flushed in between) for an `ncu` capture.
    ncu --set full --import-source on --clock-control none -k regex:'post_physics|reset_observe' --launch-skip 8 \
        --launch-count 2 -o gpurun_out/x python tools/ncu_step.py 8192 [config3]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from bench import make_cfg
from ti5_isaacgym_b200.envs import T1DHStandEnv
from ti5_isaacgym_b200.sim.synthetic import SimParams, fill_synthetic_state, synthetic_actions

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cfg = make_cfg(N)
if len(sys.argv) > 2 and sys.argv[2] == "config3":
    cfg.terrain.mesh_type = "trimesh"
    cfg.terrain.measure_heights = True
    cfg.env.num_privileged_obs = 3 * (73 + 187)
    cfg.domain_rand.push_robots = True
env = T1DHStandEnv(cfg, SimParams(dt=cfg.sim.dt), 1, "cuda:0", True, use_cuda_graph=False, materialize_obs=False)
gen = torch.Generator(device="cuda").manual_seed(1)
fill_synthetic_state(env.gym.tensors, env.env_origins, gen, base_contact_rate=0.05 if len(sys.argv) > 2 else 0.01)
env.reset()
env.episode_length_buf = torch.randint(1, 2000, (N,), generator=gen, device="cuda")
act = synthetic_actions(N, gen, "cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(int(os.environ.get("STEPS", 8))):
    flush.fill_(i)
    env.step(act)
# the GAE kernels on a 24-step rollout of the same size (rs:97-119)
from ti5_isaacgym_b200.algo.rollout_storage import gae_returns_
rew, val = torch.randn(24, N, 1, device="cuda"), torch.randn(24, N, 1, device="cuda")
done, last = (torch.rand(24, N, 1, device="cuda") < 0.02).byte(), torch.randn(N, 1, device="cuda")
ret, adv = torch.empty_like(rew), torch.empty_like(rew)
for i in range(2):
    flush.fill_(i)
    gae_returns_(rew, val, done, last, ret, adv, 0.994, 0.9)
torch.cuda.synchronize()
print("done", env.launches_per_step, "launches per step")
