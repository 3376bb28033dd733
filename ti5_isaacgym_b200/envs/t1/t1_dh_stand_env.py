"""`T1DHStandEnv`: the `t1_dh_stand` task (reference: humanoid/envs/t1/t1_dh_stand_env.py).

Everything task-specific of the reference — gait phase and stance mask (t1:80-107), gait-schedule
command resampling (t1:109-177), push / external-force windows (t1:193-247), the reference pose
(t1:250-274), the 47-dim observation and 73-dim privileged frames with lagged proprioception and
noise (t1:368-481), the T1 reset (t1:483-559) and the reward terms (t1:572-946) — is compiled into
the fused kernels `LeggedRobot.step` launches; this class contributes the robot constants and the
helpers that expose task quantities with the reference's names.
"""
import torch

from ..base.legged_robot import LeggedRobot
from .t1_robot import robot_constants


class T1DHStandEnv(LeggedRobot):
    def __init__(self, cfg, sim_params, physics_engine, sim_device, headless, **kw):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless, **kw)

    def _robot_constants(self):
        return robot_constants(self.cfg)

    # ---- task quantities recomputed on demand with plain torch (inspection only; not on the hot path)
    def _stand_command(self):
        return torch.norm(self.commands[:, :3], dim=1) <= self.cfg.commands.stand_com_threshold

    def _get_phase(self):
        """t1:80-92 without the side effect (the kernels already applied it this step)."""
        stand = self._stand_command()
        cyc = self.cfg.rewards.cycle_time
        return ((self.phase_length_buf * self.dt / cyc) % 1.0 + self.gait_start) * (~stand)

    def _get_gait_phase(self):
        """t1:95-107."""
        s = torch.sin(2 * torch.pi * self._get_phase())
        m = torch.zeros((self.num_envs, 2), device=self.device)
        m[:, 0] = s >= 0
        m[:, 1] = s < 0
        m[torch.abs(s) < 0.1] = 1
        return m
