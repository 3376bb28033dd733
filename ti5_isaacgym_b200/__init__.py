"""ti5_isaacgym_b200 — B200-native per-step environment math for the `t1_dh_stand` task.

Only the hot path of Robotics-Engineer-khy/ti5_isaacgym lives here:
hand-written sm_100a CUDA kernels behind a C ABI (`csrc/`, `include/ti5_step.h`) and the
host-side mirror of the reference's `LeggedRobot` / `T1DHStandEnv` / `task_registry` /
`RolloutStorage.compute_returns` interfaces that call them.
"""
import os

__version__ = "0.1.0"
PACKAGE_DIR = os.path.dirname(os.path.realpath(__file__))
LEGGED_GYM_ROOT_DIR = os.path.dirname(PACKAGE_DIR)
LEGGED_GYM_ENVS_DIR = os.path.join(PACKAGE_DIR, "envs")
