"""`T1DHStandEnv`: the `t1_dh_stand` task (reference: humanoid/envs/t1/t1_dh_stand_env.py).

The per-env arithmetic of the reference's task class — gait phase and stance mask (t1:80-107), gait-schedule
command resampling (t1:109-177), push / external-force draws (t1:217-247), the reference pose (t1:250-274), the
47-dim observation and 73-dim privileged frames with lagged proprioception and noise (t1:368-481), the T1 reset
(t1:483-559) and the reward terms (t1:572-946) — is compiled into the kernels `LeggedRobot.step` launches.  What stays
on the host is what the reference's task class does on the host: the integer schedule of the disturbance windows
(t1:193-215) and the hand-over of pushes and external forces to the simulator (t1:230, 247), `step`'s optional
reference-action offset (t1:360-366), and the robot constants.
"""
import torch

from ... import _lib
from ..base.legged_robot import LeggedRobot
from .t1_robot import robot_constants

C = _lib.CONSTS


class T1DHStandEnv(LeggedRobot):
    def __init__(self, cfg, sim_params, physics_engine, sim_device, headless, **kw):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless, **kw)

    def _robot_constants(self):
        return robot_constants(self.cfg)

    def step(self, actions):
        """t1:360-366: with `env.use_ref_actions` the policy's output is an offset on the reference action of the gait
        (added in place, like the reference does)."""
        if getattr(self.cfg.env, "use_ref_actions", False):
            actions += self.ref_action
        return super().step(actions)

    # ------------------------------------------------------------------ disturbances: schedule and simulator hand-over
    def _disturbance_windows(self, counter):
        """t1:193-215: the push / external-force window predicates of the step whose `common_step_counter` (after
        lr:471) is `counter` — a host-side integer schedule, the same arithmetic ti5_post_physics does on the device."""
        w = self.__dict__.get("_window_consts")
        if w is None:                      # plain Python numbers: this runs once per step on the host
            p = self._params
            w = self._window_consts = (
                bool(p.flags & C["TI5_F_PUSH_ROBOTS"]), int(p.push_update_step), int(p.push_interval), list(p.push_duration)[:p.n_push_dur],
                bool(p.flags & C["TI5_F_ADD_EXT_FORCE"]), int(p.add_update_step), int(p.ext_force_interval), list(p.add_duration)[:p.n_add_dur])
        push = force = False
        if w[0]:
            push = counter % w[2] <= w[3][min(counter // w[1], len(w[3]) - 1)]
        if w[4]:
            force = counter % w[6] <= w[7][min(counter // w[5], len(w[7]) - 1)]
        return push, force

    def _notify_simulator_of_disturbances(self):
        """t1:230 / t1:247, where `_post_physics_step_callback` issues them: the pushed base velocities (written into
        `root_states` by ti5_post_physics) go back to the simulator as the whole root tensor; the external force /
        torque on the base as (N, NB, 3) tensors in env space.  Stream-ordered behind the kernel, no host wait."""
        push, force = self._disturbance_windows(self.common_step_counter + 1)
        if push:
            self.gym.set_actor_root_state_tensor(self.sim, self.root_states)
        if force:
            self.gym.apply_rigid_body_force_tensors(self.sim, self._apply_forces, self._apply_torques, 0)   # gymapi.ENV_SPACE


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
