"""Per-launch time of the substep phase graph (10 launches) with features switched off one at a time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import sys, torch, ctypes
from bench import make_cfg
from ti5_isaacgym_b200 import _lib
from ti5_isaacgym_b200.envs import T1DHStandEnv
from ti5_isaacgym_b200.sim.synthetic import SimParams, fill_synthetic_state, synthetic_actions
C = _lib.CONSTS
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cfg = make_cfg(N)
env = T1DHStandEnv(cfg, SimParams(dt=cfg.sim.dt), 1, 'cuda:0', True, use_cuda_graph=True, materialize_obs=False)
gen = torch.Generator(device='cuda').manual_seed(1)
fill_synthetic_state(env.gym.tensors, env.env_origins, gen)
env.reset()
act = synthetic_actions(N, gen, 'cuda')
for _ in range(5): env.step(act)
torch.cuda.synchronize()
base_flags = env._params.flags
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def timeit(label, flags, do_flush):
    env._params.flags = flags
    g_sub = env.capture_phase_graphs()[0]
    ev = lambda: torch.cuda.Event(enable_timing=True)
    pairs = []
    for i in range(60):
        if do_flush: flush.fill_(i)
        a, b = ev(), ev()
        a.record(); g_sub.replay(); b.record()
        pairs.append((a, b))
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in pairs[10:])
    print("%-34s flush=%d  median %.2f us per 10 launches  (%.2f us/launch)" % (label, do_flush, t[len(t)//2]*1e3, t[len(t)//2]*100))
    env._params.flags = base_flags
for fl in (1, 0):
    timeit("all features", base_flags, fl)
    timeit("- imu lag push", base_flags & ~C["TI5_F_ADD_IMU_LAG"], fl)
    timeit("- dof lag push", base_flags & ~C["TI5_F_ADD_DOF_LAG"], fl)
    timeit("- action lag (ring)", base_flags & ~C["TI5_F_ADD_LAG"], fl)
    timeit("- torque rng (philox)", base_flags & ~C["TI5_F_RAND_TORQUE"], fl)
    timeit("- gains/friction loads", base_flags & ~(C["TI5_F_RAND_GAINS"] | C["TI5_F_RAND_COULOMB"]), fl)
    timeit("none of the above", base_flags & ~(C["TI5_F_ADD_IMU_LAG"] | C["TI5_F_ADD_DOF_LAG"] | C["TI5_F_ADD_LAG"] | C["TI5_F_RAND_TORQUE"] | C["TI5_F_RAND_GAINS"] | C["TI5_F_RAND_COULOMB"]), fl)
