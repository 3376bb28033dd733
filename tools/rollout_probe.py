"""GPU-side probe of the rollout-storage kernels (run under ncu for profiles/): fills a T=24 rollout at N envs
and draws mini-batches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import sys
import torch
from ti5_isaacgym_b200.algo.rollout_storage import FrameLogRolloutStorage, RolloutStorage
from ti5_isaacgym_b200.envs import T1DHStandEnv, make_t1_cfg
from ti5_isaacgym_b200.sim.synthetic import SimParams, fill_synthetic_state, synthetic_actions

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T, NMB = 24, 4
cfg = make_t1_cfg()()
cfg.env.num_envs = N
cfg.terrain.mesh_type = "plane"
env = T1DHStandEnv(cfg, SimParams(dt=cfg.sim.dt), 1, "cuda:0", True, rng_mode="philox", div_mode="reciprocal", materialize_obs=False)
gen = torch.Generator(device="cuda").manual_seed(1)
fill_synthetic_state(env.gym.tensors, env.env_origins, gen)
obs, priv = env.reset()
st = FrameLogRolloutStorage(env, T)
act = synthetic_actions(N, gen, "cuda")
val, logp, sig = torch.randn(N, 1, device="cuda"), torch.randn(N, device="cuda"), torch.full((N, 12), 0.3, device="cuda")
for t in range(T):
    tr = RolloutStorage.Transition()
    tr.actions, tr.values, tr.actions_log_prob, tr.action_mean, tr.action_sigma, tr.observations = act, val, logp, act, sig, obs
    obs, priv, rew, dones, infos = env.step(act)
    st.store_step(tr, rew, dones, infos["time_outs"], 0.994)
mb = T * N // NMB
idx = torch.randperm(T * N, device="cuda", generator=gen)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)
for sort in (False, True):
    times = []
    for rep in range(3):
        for i in range(NMB):
            flush.fill_(i)
            a, b = ev(), ev()
            a.record(); out = st.gather(idx[i * mb:(i + 1) * mb], sort=sort); b.record()
            times.append((a, b)); del out
    torch.cuda.synchronize()
    print("sorted" if sort else "unsorted", "gather ms per mini-batch:", [round(a.elapsed_time(b), 4) for a, b in times])
# store launch alone
tr = RolloutStorage.Transition()
tr.actions, tr.values, tr.actions_log_prob, tr.action_mean, tr.action_sigma, tr.observations = act, val, logp, act, sig, obs
st.clear()
pairs = []
for t in range(T):
    a, b = ev(), ev()
    a.record(); st.store_step(tr, rew, dones, infos["time_outs"], 0.994); b.record()
    pairs.append((a, b))
torch.cuda.synchronize()
print("store us:", [round(1e3 * a.elapsed_time(b), 1) for a, b in pairs][-8:])
