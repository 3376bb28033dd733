"""Drive the UNMODIFIED reference (`/root/reference`) with synthetic simulator tensors and
with randomness supplied as per-env pools.  TEST INFRASTRUCTURE; only usable where the
reference tree is mounted (the build container), never on the GPU box.

How the random draws are matched: the reference draws `len(env_ids)` numbers at each call
site (data-dependent length).  Every site is identified from the caller's frame (function
name, source line, the local `env_ids`) and served from the per-env pool the oracle and
the CUDA kernels consume (`oracle.t1_oracle.rng_pool_shapes`), gathered at `env_ids`.
"""
import contextlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("TI5_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "humanoid"))


def import_reference():
    """Put the shim and the reference on sys.path; returns the `humanoid.envs` module."""
    for p in (REFERENCE_ROOT, os.path.join(HERE, "shim")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import humanoid.envs as envs        # noqa: F401  (must precede humanoid.utils: circular import)
    return envs


def _fake_terrain_class():
    """The reference builds `Terrain(cfg.terrain, num_envs)` from isaacgym.terrain_utils (absent);
    both sides use the product's synthetic stand-in so they see the same height field."""
    sys.path.insert(0, os.path.dirname(HERE))
    from ti5_isaacgym_b200.sim.synthetic import SyntheticTerrain
    return SyntheticTerrain


# site tables: (function name, line) -> how to serve the draw
_RAND_FLOAT_SITES = {
    ("_compute_torques", 1071): ("torque", None),
    ("randomize_dof_props", 737): ("dr", 0), ("randomize_dof_props", 741): ("dr", 1),
    ("randomize_dof_props", 746): ("dr", 2), ("randomize_dof_props", 747): ("dr", 3),
    ("randomize_dof_props", 752): ("dr", 4), ("randomize_dof_props", 753): ("dr", 5),
    ("randomize_dof_props", 780): ("dr", 6),
    ("randomize_dof_props", 763): ("dr_joint", 0), ("randomize_dof_props", 773): ("dr_joint", 1),
    ("_reset_dofs", 1084): ("dofs", None),
    ("_reset_root_states", 1105): ("root_xy", None), ("_reset_root_states", 1108): ("root_xy", None),
    ("_resample_walk_sagittal_command", 147): ("cmd", 0),
    ("_resample_walk_lateral_command", 156): ("cmd", 1),
    ("_resample_rotate_command", 168): ("cmd", 2),
    ("_resample_walk_omnidirectional_command", 171): ("cmd", 0),
    ("_resample_walk_omnidirectional_command", 172): ("cmd", 1),
    ("_resample_walk_omnidirectional_command", 176): ("cmd", 2),
    # heading mode (t1:165-174): the third draw of the gait slot is the heading target
    ("_resample_rotate_command", 166): ("cmd", 2), ("_resample_walk_omnidirectional_command", 174): ("cmd", 2),
    ("generate_gait_time", 116): ("gait_time", None),
    ("_push_robots", 223): ("push", (0, 2)), ("_push_robots", 225): ("push", (2, 5)),
    ("_add_ext_force", 237): ("ext", (0, 1)), ("_add_ext_force", 238): ("ext", (1, 2)),
    ("_add_ext_force", 239): ("ext", (2, 3)), ("_add_ext_force", 241): ("ext", (3, 6)),
}
_RANDINT_SITES = {("randomize_lag_props", 608): 0, ("randomize_lag_props", 618): 1, ("randomize_lag_props", 628): 2}
_RANDINT_PV_SITES = {("randomize_lag_props", 639): 0, ("randomize_lag_props", 646): 1}       # position / velocity lag at a reset
# per-step re-draws over all envs: row of the `lag_step` pool behind the DEC substep rows (DOF, IMU, position, velocity)
_RANDINT_STEP_SITES = {("compute_observations", 409): 0, ("compute_observations", 438): 1,
                       ("compute_observations", 418): 2, ("compute_observations", 426): 3}


class ReferenceDriver:
    """Owns one reference `T1DHStandEnv` on CPU over the fake gym."""

    def __init__(self, num_envs, mesh_type="plane", cfg_edit=None, seed=0, device="cpu"):
        envs = import_reference()
        from isaacgym import gymapi
        import humanoid.envs.t1.t1_dh_stand_env as t1_mod
        import humanoid.envs.base.legged_robot as lr_mod
        self.t1_mod, self.lr_mod = t1_mod, lr_mod
        t1_mod.Terrain = lr_mod.Terrain = _fake_terrain_class()
        cfg = envs.DHT1StandCfg()
        cfg.env.num_envs = num_envs
        cfg.terrain.mesh_type = mesh_type
        if cfg_edit is not None:
            cfg_edit(cfg)
        torch.manual_seed(seed)
        sp = gymapi.SimParams(dt=cfg.sim.dt, use_gpu_pipeline=(device != "cpu"))
        self.cfg = cfg
        self.env = envs.T1DHStandEnv(cfg, sp, gymapi.SIM_PHYSX, device, True)
        self.pools = None
        self.substep = 0
        self._orig = dict(randint=torch.randint, rand_like=torch.rand_like, randint_like=torch.randint_like)

    # -- simulator tensors --------------------------------------------------------------
    @property
    def sim(self):
        from types import SimpleNamespace
        t = self.env.gym.tensors
        return SimpleNamespace(root_states=t["root"], dof_state=t["dof"], contact_forces=t["contact"],
                               rigid_state=t["rigid"])

    # -- RNG-as-input ----------------------------------------------------------------------
    def _frames(self):
        f = sys._getframe(2)
        out = []
        while f is not None and len(out) < 12:
            out.append(f)
            f = f.f_back
        return out

    def _rand_float(self, lower, upper, shape, device):
        fr = self._frames()
        site = (fr[0].f_code.co_name, fr[0].f_lineno)
        name, sel = _RAND_FLOAT_SITES[site]
        loc = fr[0].f_locals
        P = self.pools[name]
        if name == "torque":
            u = P[self.substep]
            self.substep += 1
        elif name == "dr":
            ids = loc["env_ids"]
            u = P[ids, 6, loc["i"]:loc["i"] + 1] if sel == 6 else P[ids, sel]
        elif name == "dr_joint":
            u = P[loc["env_ids"], sel:sel + 1]
        elif name in ("dofs", "root_xy"):
            u = P[loc["env_ids"]]
        elif name == "cmd":
            in_reset = any(f.f_code.co_name == "reset_idx" for f in fr)
            gait = next(f.f_locals["i"] for f in fr if f.f_code.co_name == "_resample_commands")
            u = P[1 if in_reset else 0, gait][loc["env_ids"], sel:sel + 1]
        elif name == "gait_time":
            u = P[loc["envs"], loc["i"]:loc["i"] + 1]
        else:
            u = P[:, sel[0]:sel[1]]
        assert tuple(u.shape) == tuple(shape), (site, u.shape, shape)
        return (upper - lower) * u + lower

    def _randint(self, *args, **kw):
        fr = self._frames()
        site = (fr[0].f_code.co_name, fr[0].f_lineno)
        if site in _RANDINT_SITES:
            return self.pools["lag_idx"][fr[0].f_locals["env_ids"], _RANDINT_SITES[site]].clone()
        if site in _RANDINT_PV_SITES:
            return self.pools["lag_idx_pv"][fr[0].f_locals["env_ids"], _RANDINT_PV_SITES[site]].clone()
        if site in _RANDINT_STEP_SITES:
            dec = self.cfg.control.decimation
            return self.pools["lag_step"][dec + _RANDINT_STEP_SITES[site]].clone()
        if site == ("_compute_torques", 1039):          # the action lag, re-drawn every substep (lr:1039-1043)
            k = self.lag_substep
            self.lag_substep += 1
            return self.pools["lag_step"][k].clone()
        if site == ("reset_idx", 523):
            return self.pools["gait_start"][fr[0].f_locals["env_ids"]].clone()
        return self._orig["randint"](*args, **kw)

    def _rand_like(self, t, **kw):
        fr = self._frames()
        if fr[0].f_code.co_name == "compute_observations":
            return self.pools["noise"].clone()
        return self._orig["rand_like"](t, **kw)

    def _randint_like(self, t, *a, **kw):
        fr = self._frames()
        if fr[0].f_code.co_name == "_update_terrain_curriculum":
            ids = fr[0].f_locals["env_ids"]
            return self.pools["terrain_level"][ids] % a[0]
        return self._orig["randint_like"](t, *a, **kw)

    @contextlib.contextmanager
    def pooled_rng(self, pools):
        self.pools, self.substep, self.lag_substep = pools, 0, 0
        saved = (self.t1_mod.torch_rand_float, self.lr_mod.torch_rand_float)
        self.t1_mod.torch_rand_float = self._rand_float
        self.lr_mod.torch_rand_float = self._rand_float
        torch.randint, torch.rand_like, torch.randint_like = self._randint, self._rand_like, self._randint_like
        try:
            yield
        finally:
            self.t1_mod.torch_rand_float, self.lr_mod.torch_rand_float = saved
            torch.randint, torch.rand_like = self._orig["randint"], self._orig["rand_like"]
            torch.randint_like = self._orig["randint_like"]
            self.pools = None

    def step(self, actions, pools):
        with self.pooled_rng(pools):
            return self.env.step(actions.clone())


O_OPTIONAL = ("last_lag_timestep last_dof_lag_timestep last_imu_lag_timestep dof_pos_lag_buffer dof_vel_lag_buffer dof_pos_lag_timestep dof_vel_lag_timestep last_dof_pos_lag_timestep last_dof_vel_lag_timestep joint_friction_coeffs joint_damping_coeffs").split()


def adopt_reference_state(S, env):
    """Copy every piece of persistent per-env state of a reference env into an oracle state
    (construction-time random draws included), so both continue from the same point."""
    same = ("torques actions last_actions last_last_actions last_dof_vel last_root_vel commands "
            "feet_air_time feet_height last_contacts base_quat base_lin_vel base_ang_vel "
            "projected_gravity base_euler_xyz feet_euler_xyz rand_push_force rand_push_torque "
            "ext_forces ext_torques ref_dof_pos ref_action gait_time gait_start env_frictions body_mass env_origins "
            "torque_multi motor_offsets randomized_p_gains randomized_d_gains randomized_joint_coulomb "
            "randomized_joint_viscous joint_armatures lag_buffer dof_lag_buffer imu_lag_buffer lag_timestep "
            "dof_lag_timestep imu_lag_timestep episode_length_buf phase_length_buf rew_buf reset_buf "
            "time_out_buf " + " ".join(O_OPTIONAL)).split()
    for name in same:
        if hasattr(env, name):          # the reference allocates some buffers only when their option is on (lr:251-349)
            setattr(S, name, getattr(env, name).clone())
    S.last_feet_z = env.last_feet_z if isinstance(env.last_feet_z, int) else env.last_feet_z.clone()
    S.contact_filt = getattr(env, "contact_filt", S.contact_filt).clone()
    S.common_step_counter = env.common_step_counter
    S.is_first_add_force, S.is_first_push = env.is_first_add_force, env.is_first_push
    S.command_ranges = {k: list(v) for k, v in env.command_ranges.items()}
    S.obs_history = torch.stack(list(env.obs_history), 0).clone()
    S.critic_history = torch.stack(list(env.critic_history), 0).clone()
    S.episode_sums = {k: v.clone() for k, v in env.episode_sums.items()}
    S.measured_heights = env.measured_heights if isinstance(env.measured_heights, int) else env.measured_heights.clone()
    if hasattr(env, "terrain_levels"):
        S.terrain_levels, S.terrain_types = env.terrain_levels.clone(), env.terrain_types.clone()
    S.extras = dict(env.extras)
    return S
