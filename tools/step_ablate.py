"""Whole-step and per-family device time (2-launch fused step, L2 flushed) with features switched off one at a time:
what the step's time is sensitive to.  Flags are flipped in the bound parameter block, no rebuild.
    python tools/step_ablate.py [envs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from bench import StepLoop, kernel_times
from ti5_isaacgym_b200 import _lib

C = _lib.CONSTS
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
loop = StepLoop(N, "cuda:0", 2, 66)
env = loop.env
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(24):
    loop.one()
base = env._params.flags
F = lambda *names: sum(C[n] for n in names)
cases = [("all features", 0),
         ("- torque rng (10 Philox / item)", F("TI5_F_RAND_TORQUE")),
         ("- obs noise (12 Philox / env)", F("TI5_F_ADD_NOISE")),
         ("- action lag ring", F("TI5_F_ADD_LAG")),
         ("- dof lag ring", F("TI5_F_ADD_DOF_LAG")),
         ("- imu lag ring", F("TI5_F_ADD_IMU_LAG")),
         ("- gains / friction arrays", F("TI5_F_RAND_GAINS", "TI5_F_RAND_COULOMB")),
         ("- all three lag rings + rng", F("TI5_F_ADD_LAG", "TI5_F_ADD_DOF_LAG", "TI5_F_ADD_IMU_LAG", "TI5_F_RAND_TORQUE", "TI5_F_ADD_NOISE"))]
for rep in range(2):
    for label, off in cases:
        env._params.flags = base & ~off
        env._bind_buffers()
        env._drop_graphs()
        for _ in range(6):
            loop.one()
        ms, _ = loop.timed(96, flush)
        kt = kernel_times(env, loop.actions, steps=24)
        print("%-36s step %.2f us   %s" % (label, ms / 96 * 1e3, {k: round(v["phase_ms"] * 1e3, 1) for k, v in kt.items()}), flush=True)
    env._params.flags = base
