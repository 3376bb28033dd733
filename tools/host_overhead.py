"""Host-side cost of env.step() (enqueue only, no sync) and where it goes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cProfile, pstats, time, torch
from ti5_isaacgym_b200.envs import T1DHStandEnv, make_t1_cfg
from ti5_isaacgym_b200.sim.synthetic import SimParams, fill_synthetic_state, synthetic_actions
N = 8192
cfg = make_t1_cfg()(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "plane"
env = T1DHStandEnv(cfg, SimParams(dt=cfg.sim.dt), 1, "cuda:0", True, rng_mode="philox", div_mode="reciprocal", materialize_obs=False)
gen = torch.Generator(device="cuda").manual_seed(1)
fill_synthetic_state(env.gym.tensors, env.env_origins, gen)
env.reset()
act = synthetic_actions(N, gen, "cuda")
h_act = act.cpu().pin_memory()
for a in (act, h_act):
    for _ in range(50): env.step(a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000): env.step(a)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("actions on", a.device, ": host enqueue %.1f us/step, incl. drain %.1f us/step" % ((t1 - t0) / 2000 * 1e6, (t2 - t0) / 2000 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): env.step(h_act)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
