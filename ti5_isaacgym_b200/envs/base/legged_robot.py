"""`LeggedRobot`: host-side mirror of `humanoid/envs/base/legged_robot.py` for the hot path.

Same constructor, same `VecEnv` methods (`step / reset / get_observations / get_privileged_
observations`), same public attributes — but every per-step torch-op chain of the
reference is one call into libti5step.so (sm_100a CUDA, `include/ti5_step.h`):

    step()                 lr:387-448    -> ti5_begin_step, 10 x ti5_substep, ti5_post_physics,
                                            ti5_reset_observe        (optionally one CUDA graph)
    _compute_torques()     lr:1019-1074  -> ti5_torque_substep
    post_physics_step()    lr:458-506    -> ti5_post_physics + ti5_reset_observe
    reset_idx()            lr:520-602    -> ti5_reset_bookkeeping + ti5_reset_scatter
    _get_heights()         lr:1551-1587  -> ti5_sample_heights

State lives in a handful of torch allocations laid out for the kernels (SoA per-env rows, slot-major
lag rings, mirrored observation rings); the reference-shaped views (`lag_buffer`,
`obs_history`, ...) are materialised on demand for tests and tools.  Asset / PhysX set-up
(lr:1239-1417) is out of scope: the simulator handle is anything with the gym tensor API.
"""
import ctypes
import os
from types import SimpleNamespace

import numpy as np
import torch

from ... import _lib
from ...sim.synthetic import SyntheticTerrain
from ...utils.helpers import class_to_dict
from .base_task import BaseTask
from .step_params import TERM_NAMES, build_params, reward_scales

C = _lib.CONSTS
STEP_INDEX0 = 4          # ring pushes "before time zero": STEP_INDEX0 * decimation >= longest lag ring


class EpisodeInfo(dict):
    """`extras["episode"]` (t1:530-538) backed by one row of the device-side log ring; entries are
    created on first access so a step that nobody inspects costs no extra launches."""

    def __init__(self, row, names, with_terrain, with_curriculum):
        super().__init__()
        self._row, self._names, self._wt, self._wc, self._done = row, names, with_terrain, with_curriculum, False

    def _fill(self):
        if self._done:
            return
        self._done = True
        for name in self._names:
            dict.__setitem__(self, "rew_" + name, self._row[TERM_NAMES.index(name)])
        if self._wt:
            dict.__setitem__(self, "terrain_level", self._row[C["TI5_NUM_TERMS"]])
        if self._wc:
            dict.__setitem__(self, "max_command_x", float(self._row[C["TI5_NUM_TERMS"] + 1]))

    def __getitem__(self, k):
        self._fill()
        return dict.__getitem__(self, k)

    def __iter__(self):
        self._fill()
        return dict.__iter__(self)

    def __len__(self):
        self._fill()
        return dict.__len__(self)

    def __contains__(self, k):
        self._fill()
        return dict.__contains__(self, k)

    def keys(self):
        self._fill()
        return dict.keys(self)

    def items(self):
        self._fill()
        return dict.items(self)

    def values(self):
        self._fill()
        return dict.values(self)


class LeggedRobot(BaseTask):
    def __init__(self, cfg, sim_params, physics_engine, sim_device, headless, gym=None, rng_mode="philox",
                 div_mode="reciprocal", use_cuda_graph=True, materialize_obs=True, seed=None, chain_launches=None,
                 fused_step=None):
        """Args as the reference (lr:57).  Extra keyword options:
        rng_mode        "philox" (in-kernel Philox4x32-10) or "pools" (uniforms supplied per step via
                        `set_rng_pools`, the parity mode of SURVEY.md section 7)
        div_mode        "reciprocal" = torch-on-GPU rounding of tensor/scalar, "ieee" = torch-on-CPU
        use_cuda_graph  capture the whole step into one CUDA graph (philox mode, synthetic sim only)
        materialize_obs True (default): `step` returns FRESH contiguous (N, H*47) / (N, CH*P) tensors, as the reference
                        does (t1:477-481) — a caller may hold them across later steps, which the reference's runner does
                        (dh_ppo.py:88 -> rs:62).  False: zero-copy views into the history rings, valid ONLY until the
                        next `step()` (+25 KB/env/step saved); for callers that consume the observation at once, or
                        that use `FrameLogRolloutStorage` (`task_registry.make_alg_runner` switches the env to views
                        when it installs that storage)
        chain_launches  launch the kernels of a fused step (no simulator in between) as programmatic dependents of one
                        another: each becomes resident while its predecessor still runs and waits for it only where
                        it needs its results (default on; env var TI5_CHAIN=0 turns it off)
        fused_step      with no simulator between the substeps, run the action clip, the DEC substeps and the post-physics
                        phase as ONE launch (ti5_fused_step) followed by ti5_reset_observe: 2 launches per step instead of
                        12, bit-identical results (default on; env var TI5_FUSED=0 selects the 12-launch sequence)"""
        self.cfg = cfg
        self.sim_params = sim_params
        self.height_samples = None
        self.debug_viz = False
        self.init_done = False
        self._lib = _lib.load_library()
        self._rng_mode = C["TI5_RNG_PHILOX"] if rng_mode == "philox" else C["TI5_RNG_POOLS"]
        self._div_mode = C["TI5_DIV_RECIPROCAL"] if div_mode == "reciprocal" else C["TI5_DIV_IEEE"]
        self._use_graph = bool(use_cuda_graph) and rng_mode == "philox"
        self._materialize = bool(materialize_obs)
        self._chain_launches = (os.environ.get("TI5_CHAIN", "1") != "0") if chain_launches is None else bool(chain_launches)
        self._fused_step = (os.environ.get("TI5_FUSED", "1") != "0") if fused_step is None else bool(fused_step)
        self._seed = int(getattr(cfg, "seed", 0) if seed is None else seed)
        self._parse_cfg(self.cfg)
        super().__init__(self.cfg, sim_params, physics_engine, sim_device, headless, gym=gym)
        self._init_buffers()
        self._prepare_reward_function()
        self.init_done = True

    # ------------------------------------------------------------------ configuration (lr:94-113)
    def _parse_cfg(self, cfg):
        self.dt = cfg.control.decimation * self.sim_params.dt
        self.obs_scales = cfg.normalization.obs_scales
        self.reward_scales = class_to_dict(cfg.rewards.scales)
        if cfg.terrain.mesh_type not in ("heightfield", "trimesh"):
            cfg.terrain.curriculum = False
        self.max_episode_length_s = cfg.env.episode_length_s
        self.max_episode_length = np.ceil(self.max_episode_length_s / self.dt)
        cfg.domain_rand.push_interval = np.ceil(cfg.domain_rand.push_interval_s / self.dt)
        cfg.domain_rand.ext_force_interval = np.ceil(cfg.domain_rand.ext_force_interval_s / self.dt)

    def _robot_constants(self):
        raise NotImplementedError("the task class supplies the robot model constants")

    # ------------------------------------------------------------------ simulator handle (lr:1239-1257)
    def create_sim(self):
        self.up_axis_idx = 2
        self.sim = self.gym
        mesh = self.cfg.terrain.mesh_type
        if mesh in ("heightfield", "trimesh"):
            self.terrain = getattr(self.cfg.terrain, "terrain_object", None) or SyntheticTerrain(self.cfg.terrain, self.num_envs)
            self.height_samples = torch.as_tensor(self.terrain.heightsamples).view(
                self.terrain.tot_rows, self.terrain.tot_cols).to(self.device)
        elif mesh not in ("plane", None):
            raise ValueError("Terrain mesh type not recognised. Allowed types are [None, plane, heightfield, trimesh]")
        self.robot = self._robot_constants()
        self.num_dof = self.num_dofs = self.robot.num_dof
        self.num_bodies = self.robot.num_bodies
        self.dof_names = list(self.robot.dof_names)
        dev = self.device
        as_idx = lambda v: torch.tensor(v, dtype=torch.long, device=dev)
        self.feet_indices, self.knee_indices = as_idx(self.robot.feet_indices), as_idx(self.robot.knee_indices)
        self.penalised_contact_indices = as_idx(self.robot.penalised_contact_indices)
        self.termination_contact_indices = as_idx(self.robot.termination_contact_indices)
        self._get_env_origins()
        self.envs = list(range(self.num_envs))
        self.actor_handles = [0] * self.num_envs

    def _get_env_origins(self):
        """lr:1477-1512: terrain platform origins, or a square grid on a plane."""
        dev, N, t = self.device, self.num_envs, self.cfg.terrain
        self.env_origins = torch.zeros(N, 3, device=dev)
        if t.mesh_type in ("heightfield", "trimesh"):
            self.custom_origins = True
            max_init = t.max_init_terrain_level if t.curriculum else t.num_rows - 1
            self.terrain_levels = torch.randint(0, max_init + 1, (N,), device=dev)
            self.terrain_types = torch.div(torch.arange(N, device=dev), (N / t.num_cols), rounding_mode="floor").to(torch.long)
            self.max_terrain_level = t.num_rows
            self.terrain_origins = torch.from_numpy(np.asarray(self.terrain.env_origins)).to(dev).to(torch.float).contiguous()
            self.env_origins[:] = self.terrain_origins[self.terrain_levels, self.terrain_types]
        else:
            self.custom_origins = False
            cols = np.floor(np.sqrt(N))
            rows = np.ceil(N / cols)
            xx, yy = torch.meshgrid(torch.arange(rows), torch.arange(cols), indexing="ij")
            spacing = self.cfg.env.env_spacing
            self.env_origins[:, 0] = spacing * xx.flatten()[:N].to(dev)
            self.env_origins[:, 1] = spacing * yy.flatten()[:N].to(dev)

    # ------------------------------------------------------------------ buffers (lr:116-349, base_task.py:55-74)
    def _init_buffers(self):
        cfg, dev, N, D, NB = self.cfg, self.device, self.num_envs, self.num_dof, self.num_bodies
        dr = cfg.domain_rand
        f32 = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        i64 = lambda *s: torch.zeros(*s, dtype=torch.int64, device=dev)
        u8 = lambda *s: torch.zeros(*s, dtype=torch.bool, device=dev)
        gym = self.gym
        self.root_states = gym.acquire_actor_root_state_tensor(self.sim)
        self.dof_state = gym.acquire_dof_state_tensor(self.sim)
        self.contact_forces = gym.acquire_net_contact_force_tensor(self.sim).view(N, -1, 3)
        self.rigid_state = gym.acquire_rigid_body_state_tensor(self.sim).view(N, NB, 13)
        for t in (self.root_states, self.dof_state, self.contact_forces, self.rigid_state):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        self.dof_pos = self.dof_state.view(N, D, 2)[..., 0]
        self.dof_vel = self.dof_state.view(N, D, 2)[..., 1]
        terrain = getattr(self, "terrain", None)
        hshape = tuple(self.height_samples.shape) if self.height_samples is not None else (0, 0)
        self._params = build_params(cfg, self.sim_params.dt, self.robot, terrain, hshape, self._div_mode,
                                    self._rng_mode, self._seed)
        p = self._params
        H, CH, K, P = p.frame_stack, p.c_frame_stack, p.num_single_obs, p.priv_frame
        self.common_step_counter = 0
        self.extras = {}
        self.gravity_vec = torch.tensor([0., 0., -1.], device=dev).repeat((N, 1))
        self.forward_vec = torch.tensor([1., 0., 0.], device=dev).repeat((N, 1))
        self.noise_scale_vec = torch.tensor(list(p.noise_vec)[:K], device=dev)
        self.add_noise = cfg.noise.add_noise
        self.commands_scale = torch.tensor(list(p.cmd_scale), device=dev)
        self.default_dof_pos = torch.tensor(list(p.default_dof_pos), device=dev).unsqueeze(0)
        self.default_joint_pd_target = self.default_dof_pos.clone()
        self.p_gains, self.d_gains = torch.tensor(list(p.p_gains), device=dev), torch.tensor(list(p.d_gains), device=dev)
        self.torque_limits = torch.tensor(list(p.torque_limits), device=dev)
        self.dof_vel_limits = torch.tensor(list(p.dof_vel_limits), device=dev)
        self.base_init_state = torch.tensor(list(p.base_init_state), device=dev)
        # per-env state, names as in the reference
        self.torques, self.torque_multi = f32(N, D), torch.ones(N, D, device=dev)
        self.torques_substeps = f32(p.decimation, N, D)      # lr:401-403: the torques of every substep of the last step
        self.actions, self.last_actions, self.last_last_actions = f32(N, D), f32(N, D), f32(N, D)
        self._actions_in = f32(N, D)
        self.last_dof_vel, self.last_root_vel = f32(N, D), f32(N, 6)
        self.randomized_p_gains, self.randomized_d_gains, self.motor_offsets = f32(N, D), f32(N, D), f32(N, D)
        self.randomized_joint_coulomb, self.randomized_joint_viscous, self.joint_armatures = f32(N, D), f32(N, D), f32(N, D)
        self.commands = f32(N, cfg.commands.num_commands)
        assert cfg.commands.num_commands == 4, "commands are read as one float4 per env"
        self._episode_length_buf, self.phase_length_buf = i64(N), i64(N)
        self.gait_time = torch.zeros(N, len(cfg.commands.gait), dtype=torch.int32, device=dev)
        self.gait_start = f32(N)
        self.feet_air_time, self.feet_height, self.last_feet_z = f32(N, 2), f32(N, 2), f32(N, 2)
        self.last_contacts, self.contact_filt = u8(N, 2), u8(N, 2)
        self.base_quat = f32(N, 4)
        self.base_quat[:, 3] = 1
        self.base_lin_vel, self.base_ang_vel, self.projected_gravity = f32(N, 3), f32(N, 3), f32(N, 3)
        self.projected_gravity[:, 2] = -1
        self.base_euler_xyz, self.feet_euler_xyz = f32(N, 3), f32(N, 2, 3)
        self.ref_dof_pos, self.ref_action = f32(N, D), f32(N, D)
        self.ext_forces, self.ext_torques = f32(N, 3), f32(N, 3)
        self.rand_push_force, self.rand_push_torque = f32(N, 3), f32(N, 3)
        # what apply_rigid_body_force_tensors takes (t1:234-247): (N, NB, 3) force / torque tensors, zero except the
        # base rows, which ti5_post_physics writes in place (Ti5Params.applied_stride)
        self._apply_forces, self._apply_torques = f32(N, NB, 3), f32(N, NB, 3)
        self.applied_force, self.applied_torque = self._apply_forces[:, 0, :], self._apply_torques[:, 0, :]
        p.applied_stride = 3 * NB
        self._dof_props = f32(N, D, 3)          # lr:915-939 as one dense tensor, rows in reset_ids order
        self.env_frictions, self.body_mass = f32(N, 1), f32(N, 1)
        # the per-step scalar outputs share one allocation, so a host consumer needs a single D2H copy:
        # [rew_buf f32 (4N bytes) | reset_buf bool (N) | extras["time_outs"] bool (N)]
        self.step_outputs_packed = torch.zeros(6 * N, dtype=torch.uint8, device=dev)
        self.rew_buf = self.step_outputs_packed[:4 * N].view(torch.float32)
        self.reset_buf = self.step_outputs_packed[4 * N:5 * N].view(torch.bool)
        self._time_outs_latched = self.step_outputs_packed[5 * N:].view(torch.bool)
        self.time_out_buf = u8(N)
        self.reset_buf[:] = True
        self._episode_sums = f32(C["TI5_NUM_TERMS"], N)
        self._reward_terms = f32(C["TI5_NUM_TERMS"], N)
        self.reset_ids = torch.zeros(N, dtype=torch.int32, device=dev)
        self._reset_list = torch.zeros(N, dtype=torch.int32, device=dev)
        # kernel-side layout
        self._act_ring, self._dof_ring, self._imu_ring = f32(p.lag_len, N, D), f32(p.dof_lag_len, N, 2 * D), f32(p.imu_lag_len, N, 6)
        self._lag_timestep = torch.zeros(N, 3, dtype=torch.int32, device=dev)
        # options t1_cfg leaves off: separate joint position / velocity lags, per-step re-draws of the lag indices
        # (`last_*_lag_timestep`, lr:282-345: the range maximum until the first re-draw; two copies, see ti5_step.h)
        hi5 = torch.tensor([p.lag_range[k][1] for k in range(3)] + [p.lag_range_pv[k][1] for k in range(2)],
                           dtype=torch.int32, device=dev)
        # (allocated at full size only with one of the options on: the t1 configuration keeps its allocation pattern)
        lag_opts = sum(C[k] for k in ("TI5_F_LAG_PERSTEP", "TI5_F_DOF_LAG_PERSTEP", "TI5_F_IMU_LAG_PERSTEP", "TI5_F_POS_VEL_LAG",
                                      "TI5_F_POS_LAG_PERSTEP", "TI5_F_VEL_LAG_PERSTEP"))
        n_opt = N if p.flags & lag_opts else 1
        self._joint_coeffs = torch.ones(N if p.flags2 else 1, 2, dtype=torch.float32, device=dev)      # lr:1449-1457
        self._lag_pv = hi5[3:5].repeat(n_opt, 1).contiguous()
        self._last_lag = hi5.repeat(2, n_opt, 1).contiguous()
        self._ring_stamp = i64(N)
        self._obs_ring, self._priv_ring = f32(N, 2 * H, K), f32(N, 2 * CH, P)
        nblk = (N + 31) // 32
        self._block_counts = torch.zeros(nblk + 1, dtype=torch.int32, device=dev)
        self._block_sums = f32(nblk, C["TI5_LOG_COLS"])
        self._extras_log = f32(C["TI5_LOG_ROWS"], C["TI5_LOG_COLS"])
        self._globals = torch.zeros(ctypes.sizeof(_lib.Ti5Globals), dtype=torch.uint8, device=dev)
        # len(env_ids) of the step in progress (lr:490) where it lives: a 0-dim int32 view into the device globals
        off = _lib.Ti5Globals.n_reset.offset
        self.n_reset_device = self._globals[off:off + 4].view(torch.int32)[0]
        self._debug_ts = None        # set to a (2, 4096, 8) int64 CUDA tensor and re-bind to collect kernel probes
        self._obs_out = self._priv_out = None                     # allocated per step by _materialize_windows()
        self._frame_log = self._priv_log = self._valid_log = self._hist_valid = None     # enable_frame_log()
        self.measured_heights = f32(N, max(p.num_height_points, 1)) if p.num_height_points else 0
        self._height_points = None
        if cfg.terrain.measure_heights:
            self.height_points = self._init_height_points()
            self._height_points = self.height_points[0, :, :2].contiguous()
        if not self.custom_origins:
            self.terrain_levels, self.terrain_types = i64(N), i64(N)
            self.terrain_origins = f32(1, 1, 3)
        # construction-time randomisation (lr:1371-1373, 277-315; t1:569): plain torch, runs once
        self.command_ranges = class_to_dict(cfg.commands.ranges)
        self._randomize_initial_state()
        self._step_index = STEP_INDEX0
        self._write_globals()
        self._bind_buffers()
        self._graph = self._graph_host = None
        self._view_epoch = [STEP_INDEX0]            # mutable cell shared with the ring-view tags: [current step index]
        self._view_cache = ({}, {})                 # ring-window views by slot (`_window_views`)
        self._extras_rows = [self._extras_log[i] for i in range(C["TI5_LOG_ROWS"])]
        self._extras_args = (list(reward_scales(cfg, self.dt)), cfg.terrain.mesh_type == "trimesh", bool(cfg.commands.curriculum),
                             bool(cfg.env.send_timeouts))
        self._graphs, self._max_graphs = {}, 8      # captured steps keyed by the address of the action buffer
        self.host_actions = None                                  # enable_host_io()
        self._rng = None
        self._rng_keep = None
        self.obs_buf, self.privileged_obs_buf = self._history_views()

    def _init_height_points(self):
        """lr:1535-1549."""
        y = torch.tensor(self.cfg.terrain.measured_points_y, device=self.device)
        x = torch.tensor(self.cfg.terrain.measured_points_x, device=self.device)
        gx, gy = torch.meshgrid(x, y, indexing="ij")
        self.num_height_points = gx.numel()
        pts = torch.zeros(self.num_envs, self.num_height_points, 3, device=self.device)
        pts[:, :, 0], pts[:, :, 1] = gx.flatten(), gy.flatten()
        return pts

    def _randomize_initial_state(self):
        """The construction-time draws of the reference (torch RNG, once): actuator randomisation for
        all envs (lr:1373), lag indices (lr:277-315), friction buckets and payloads (lr:797-822, 699)."""
        dr, N, D, dev = self.cfg.domain_rand, self.num_envs, self.num_dof, self.device
        p = self._params
        u = lambda *s: torch.rand(*s, device=dev)
        if dr.randomize_torque:
            self.torque_multi[:] = p.torque_multi_w * u(N, D) + p.torque_multi_lo
        if dr.randomize_motor_offset:
            self.motor_offsets[:] = p.motor_offset_w * u(N, D) + p.motor_offset_lo
        if dr.randomize_gains:
            self.randomized_p_gains[:] = (p.kp_mult_w * u(N, D) + p.kp_mult_lo) * self.p_gains
            self.randomized_d_gains[:] = (p.kd_mult_w * u(N, D) + p.kd_mult_lo) * self.d_gains
        if dr.randomize_coulomb_friction:
            self.randomized_joint_coulomb[:] = p.coulomb_w * u(N, D) + p.coulomb_lo
            self.randomized_joint_viscous[:] = p.viscous_w * u(N, D) + p.viscous_lo
        if dr.randomize_joint_armature:
            w, lo = torch.tensor(list(p.armature_w), device=dev), torch.tensor(list(p.armature_lo), device=dev)
            self.joint_armatures[:] = w * u(N, D) + lo
        if p.flags2 & C["TI5_F2_RAND_JOINT_FRICTION"]:                        # lr:762-763
            self._joint_coeffs[:, 0] = p.joint_friction_w * u(N) + p.joint_friction_lo
        if p.flags2 & C["TI5_F2_RAND_JOINT_DAMPING"]:                         # lr:772-773
            self._joint_coeffs[:, 1] = p.joint_damping_w * u(N) + p.joint_damping_lo
        for col, (on, rnd, rng) in enumerate(((dr.add_lag, dr.randomize_lag_timesteps, dr.lag_timesteps_range),
                                              (dr.add_dof_lag, dr.randomize_dof_lag_timesteps, dr.dof_lag_timesteps_range),
                                              (dr.add_imu_lag, dr.randomize_imu_lag_timesteps, dr.imu_lag_timesteps_range))):
            if on:
                self._lag_timestep[:, col] = torch.randint(rng[0], rng[1] + 1, (N,), device=dev) if rnd else rng[1]
        if p.flags & C["TI5_F_POS_VEL_LAG"]:                                  # lr:322-349
            for col, (rnd, rng) in enumerate(((dr.randomize_dof_pos_lag_timesteps, dr.dof_pos_lag_timesteps_range),
                                              (dr.randomize_dof_vel_lag_timesteps, dr.dof_vel_lag_timesteps_range))):
                self._lag_pv[:, col] = torch.randint(rng[0], rng[1] + 1, (N,), device=dev) if rnd else rng[1]
        self.gait_start[:] = torch.randint(0, 2, (N,), device=dev) * 0.5
        if dr.randomize_friction:
            buckets = dr.friction_range[0] + (dr.friction_range[1] - dr.friction_range[0]) * u(256)
            self.env_frictions[:, 0] = buckets[torch.randint(0, 256, (N,), device=dev)]
        mass = torch.full((N,), 10.0, device=dev)
        if dr.randomize_base_mass:
            mass = mass + dr.added_mass_range[0] + (dr.added_mass_range[1] - dr.added_mass_range[0]) * u(N)
        self.body_mass[:, 0] = mass

    # ------------------------------------------------------------------ C-ABI structs
    def _write_globals(self):
        g = _lib.Ti5Globals()
        g.step_index = self._step_index
        g.common_step_offset = self.common_step_counter - self._step_index
        for par in (0, 1):
            g.is_first_add_force[par] = 1
            for i, key in enumerate(("lin_vel_x", "lin_vel_y", "ang_vel_yaw")):
                g.cmd_range[par][i][0], g.cmd_range[par][i][1] = self.command_ranges[key]
        self._globals.copy_(torch.frombuffer(bytearray(bytes(g)), dtype=torch.uint8))

    def _read_globals(self):
        raw = bytes(self._globals.cpu().numpy().tobytes())
        return _lib.Ti5Globals.from_buffer_copy(raw)

    def _bind_buffers(self):
        b = _lib.Ti5Buffers()
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None and torch.is_tensor(t) else None
        b.globals = ptr(self._globals)
        pairs = dict(
            root_states=self.root_states, dof_state=self.dof_state, contact_forces=self.contact_forces,
            rigid_state=self.rigid_state, actions=self.actions, torques=self.torques, torques_substeps=self.torques_substeps,
            torque_multi=self.torque_multi,
            p_gains_r=self.randomized_p_gains, d_gains_r=self.randomized_d_gains, motor_offsets=self.motor_offsets,
            coulomb=self.randomized_joint_coulomb, viscous=self.randomized_joint_viscous,
            joint_armatures=self.joint_armatures, act_ring=self._act_ring, dof_ring=self._dof_ring,
            imu_ring=self._imu_ring, lag_timestep=self._lag_timestep, ring_stamp=self._ring_stamp,
            lag_pv=self._lag_pv, last_lag=self._last_lag, joint_coeffs=self._joint_coeffs,
            last_actions=self.last_actions, last_last_actions=self.last_last_actions, last_dof_vel=self.last_dof_vel,
            last_root_vel=self.last_root_vel, commands=self.commands, episode_length_buf=self._episode_length_buf,
            phase_length_buf=self.phase_length_buf, gait_time=self.gait_time, gait_start=self.gait_start,
            feet_air_time=self.feet_air_time, feet_height=self.feet_height, last_feet_z=self.last_feet_z,
            last_contacts=self.last_contacts, contact_filt=self.contact_filt, base_quat=self.base_quat,
            base_lin_vel=self.base_lin_vel, base_ang_vel=self.base_ang_vel, projected_gravity=self.projected_gravity,
            base_euler_xyz=self.base_euler_xyz, feet_euler_xyz=self.feet_euler_xyz, ref_dof_pos=self.ref_dof_pos,
            ref_action=self.ref_action, ext_forces=self.ext_forces, ext_torques=self.ext_torques,
            rand_push_force=self.rand_push_force, rand_push_torque=self.rand_push_torque,
            applied_force=self.applied_force, applied_torque=self.applied_torque, env_frictions=self.env_frictions,
            body_mass=self.body_mass, env_origins=self.env_origins, terrain_levels=self.terrain_levels,
            terrain_types=self.terrain_types, terrain_origins=self.terrain_origins, height_samples=self.height_samples,
            height_points=self._height_points,
            measured_heights=self.measured_heights if torch.is_tensor(self.measured_heights) else None,
            rew_buf=self.rew_buf, reset_buf=self.reset_buf, time_out_buf=self.time_out_buf,
            time_outs_latched=self._time_outs_latched, episode_sums=self._episode_sums,
            reward_terms=self._reward_terms, reset_ids=self.reset_ids, reset_list=self._reset_list, block_counts=self._block_counts,
            block_sums=self._block_sums, extras_log=self._extras_log, obs_ring=self._obs_ring,
            priv_ring=self._priv_ring, obs_out=self._obs_out, priv_out=self._priv_out, debug_ts=self._debug_ts,
            frame_log=self._frame_log, priv_log=self._priv_log, valid_log=self._valid_log, hist_valid=self._hist_valid,
            host_out=getattr(self, "host_outputs", None))
        pairs["dof_props"] = self._dof_props
        for name, t in pairs.items():
            if t is not None and name not in ("applied_force", "applied_torque"):      # strided rows of (N, NB, 3)
                assert t.is_contiguous(), name
            setattr(b, name, ptr(t))
        self._buffers = b
        self._keepalive = pairs
        self._p_ref, self._b_ref = ctypes.byref(self._params), ctypes.byref(self._buffers)

    # ------------------------------------------------------------------ reward bookkeeping (lr:352-384)
    def _prepare_reward_function(self):
        self.reward_scales = reward_scales(self.cfg, self.dt)
        self.reward_names = [n for n in self.reward_scales if n != "termination"]
        self.episode_sums = {n: self._episode_sums[TERM_NAMES.index(n)] for n in self.reward_scales}
        self.reward_terms = {n: self._reward_terms[TERM_NAMES.index(n)] for n in self.reward_scales}

    # ------------------------------------------------------------------ properties mirroring reference attributes
    @property
    def episode_length_buf(self):
        return self._episode_length_buf

    @episode_length_buf.setter
    def episode_length_buf(self, value):
        # the PPO runner REBINDS this attribute (dh_on_policy_runner.py:101); keep the kernel's buffer
        self._episode_length_buf.copy_(torch.as_tensor(value, device=self.device).to(torch.int64))

    @property
    def lag_timestep(self):
        return self._lag_timestep[:, 0].long()

    @property
    def dof_lag_timestep(self):
        return self._lag_timestep[:, 1].long()

    @property
    def imu_lag_timestep(self):
        return self._lag_timestep[:, 2].long()

    @property
    def joint_friction_coeffs(self):
        return self._joint_coeffs[:, 0:1]

    @property
    def joint_damping_coeffs(self):
        return self._joint_coeffs[:, 1:2]

    @property
    def dof_pos_lag_timestep(self):
        return self._lag_pv[:, 0].long()

    @property
    def dof_vel_lag_timestep(self):
        return self._lag_pv[:, 1].long()

    def last_lag_timesteps(self):
        """The reference's `last_{lag,dof_lag,imu_lag,dof_pos_lag,dof_vel_lag}_timestep` (N,5) as the NEXT re-draw will see
        them: the action lag's copy is picked by the parity of the substep count, the others' by the step's."""
        pushes = self._step_index * self._params.decimation
        out = self._last_lag[(self._step_index + 1) & 1].clone()     # step s + 1 reads the copy step s wrote
        out[:, 0] = self._last_lag[pushes & 1][:, 0]
        return out.long()

    def _ring_as_shifted(self, ring, length):
        """(len, N, W) slot-major ring -> the reference's (N, W, len) shifted buffer, slot 0 newest."""
        pushes = self._step_index * self._params.decimation
        ages = torch.arange(length, device=self.device)
        idx = pushes - 1 - ages                                     # push index of each age
        rows = ring[(idx % length).clamp(min=0)]                    # (len, N, W)
        valid = (idx.view(-1, 1) >= self._ring_stamp.view(1, -1)) & (idx.view(-1, 1) >= 0)
        return (rows * valid.unsqueeze(-1)).permute(1, 2, 0).contiguous()

    @property
    def lag_buffer(self):
        return self._ring_as_shifted(self._act_ring, self._params.lag_len)

    @property
    def dof_lag_buffer(self):
        return self._ring_as_shifted(self._dof_ring, self._params.dof_lag_len)

    @property
    def imu_lag_buffer(self):
        return self._ring_as_shifted(self._imu_ring, self._params.imu_lag_len)

    @property
    def dof_pos_lag_buffer(self):
        """lr:322-326 with `add_dof_pos_vel_lag`: the position half of the DOF ring, cut to the position lag range."""
        L = self.cfg.domain_rand.dof_pos_lag_timesteps_range[1] + 1
        return self._ring_as_shifted(self._dof_ring, self._params.dof_lag_len)[:, :self.num_dof, :L]

    @property
    def dof_vel_lag_buffer(self):
        L = self.cfg.domain_rand.dof_vel_lag_timesteps_range[1] + 1
        return self._ring_as_shifted(self._dof_ring, self._params.dof_lag_len)[:, self.num_dof:, :L]

    def _history_views(self):
        """The current H-frame / CH-frame windows as (N, H*K) / (N, CH*P) views into the mirrored rings (fresh tensor
        objects; the per-step path takes them from `_window_views`)."""
        p = self._params
        H, CH, K, P = p.frame_stack, p.c_frame_stack, p.num_single_obs, p.priv_frame
        s = self._step_index
        o_off = (((s - 1) % H) + 1) * K
        c_off = (((s - 1) % CH) + 1) * P
        obs = self._obs_ring.view(self.num_envs, -1)[:, o_off:o_off + H * K]
        priv = self._priv_ring.view(self.num_envs, -1)[:, c_off:c_off + CH * P]
        return obs, priv

    def _window_views(self):
        """The same windows from a per-slot cache: the window of step s starts at ring slot s % H, so there are only H
        (CH) distinct views; building them once keeps four tensor ops out of every step."""
        p = self._params
        H, CH, K, P = p.frame_stack, p.c_frame_stack, p.num_single_obs, p.priv_frame
        s = self._step_index
        cache = self._view_cache
        ko, kc = (s - 1) % H, (s - 1) % CH
        obs = cache[0].get(ko)
        if obs is None:
            obs = cache[0][ko] = self._obs_ring.view(self.num_envs, -1)[:, (ko + 1) * K:(ko + 1) * K + H * K]
        priv = cache[1].get(kc)
        if priv is None:
            priv = cache[1][kc] = self._priv_ring.view(self.num_envs, -1)[:, (kc + 1) * P:(kc + 1) * P + CH * P]
        return obs, priv

    @property
    def obs_history(self):
        p = self._params
        return list(self._history_views()[0].reshape(self.num_envs, p.frame_stack, p.num_single_obs).unbind(1))

    @property
    def critic_history(self):
        p = self._params
        return list(self._history_views()[1].reshape(self.num_envs, p.c_frame_stack, p.priv_frame).unbind(1))

    # ------------------------------------------------------------------ frame logs (rollout storage)
    def enable_frame_log(self, num_steps):
        """Keep every appended observation / privileged frame of the last `num_steps` + H env steps, un-cleared, in
        `(N, L, K)` / `(N, L, P)` logs next to the history rings (188 + 292 B per env and step).  The rollout
        storage (`algo.rollout_storage.FrameLogRolloutStorage`) rebuilds the (N, H*K) windows of a mini-batch
        from them instead of storing (T, N, H*K) observations (rs:30-31, 62-63)."""
        p = self._params
        H, CH, K, P, N = p.frame_stack, p.c_frame_stack, p.num_single_obs, p.priv_frame, self.num_envs
        L = int(num_steps) + max(H, CH) + 2
        if p.log_len >= L:
            return self
        dev = self.device
        self._frame_log = torch.zeros(N, L, K, dtype=torch.float32, device=dev)
        self._priv_log = torch.zeros(N, L, P, dtype=torch.float32, device=dev)
        self._valid_log = torch.zeros(L, N, dtype=torch.int16, device=dev)
        self._hist_valid = torch.zeros(N, dtype=torch.int32, device=dev)
        p.log_len = L
        self._bind_buffers()
        self._drop_graphs()
        self._seed_frame_log()
        return self

    def _seed_frame_log(self):
        """The frames already in the rings become the newest rows of the logs (cleared frames are zeros there, so
        they count as valid)."""
        p = self._params
        L, s = p.log_len, self._step_index
        obs, priv = self._history_views()
        for log, win, n, w in ((self._frame_log, obs, p.frame_stack, p.num_single_obs),
                               (self._priv_log, priv, p.c_frame_stack, p.priv_frame)):
            rows = (s - n + torch.arange(n, device=self.device)) % L           # oldest .. newest = row (s - 1) % L
            log[:, rows] = win.reshape(self.num_envs, n, w)
        self._hist_valid.fill_(p.frame_stack)
        self._valid_log[(s - 1) % L] = p.frame_stack

    def frame_logs(self):
        """What the rollout storage binds: the log tensors and their geometry."""
        from types import SimpleNamespace
        p = self._params
        if not p.log_len:
            raise _lib.Ti5Error("frame logs are off: call enable_frame_log(num_steps) first")
        return SimpleNamespace(frame_stack=p.frame_stack, c_frame_stack=p.c_frame_stack, num_single_obs=p.num_single_obs,
                               priv_frame=p.priv_frame, log_len=p.log_len, frame_log=self._frame_log,
                               priv_log=self._priv_log, valid_log=self._valid_log)

    @property
    def frame_log_row(self):
        """Log row of the newest frame of the current observation."""
        return (self._step_index - 1) % self._params.log_len

    # ------------------------------------------------------------------ randomness
    def set_rng_pools(self, pools):
        """Parity mode: the uniforms of the next step (name -> tensor, `oracle.t1_oracle.rng_pool_shapes`)."""
        r = _lib.Ti5Rng()
        keep = {}
        for name, _ in _lib.Ti5Rng._fields_:
            t = pools.get(name)
            if t is None:
                continue
            want = torch.int64 if name in ("lag_idx", "gait_start", "terrain_level", "lag_idx_pv", "lag_step") else torch.float32
            t = t.to(device=self.device, dtype=want).contiguous()
            keep[name] = t
            setattr(r, name, ctypes.c_void_p(t.data_ptr()))
        self._rng, self._rng_keep = r, keep

    def _rng_ref(self):
        if self._rng_mode == C["TI5_RNG_POOLS"]:
            if self._rng is None:
                raise _lib.Ti5Error("rng_mode='pools': call set_rng_pools() before every step")
            return ctypes.byref(self._rng)
        return None

    # ------------------------------------------------------------------ the step (lr:387-448)
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chain(self, bit):
        """The *_CHAINED launch option (programmatic dependent launch) for the kernels of a step that follow one
        another on the stream with nothing in between."""
        return C[bit] if self._chain_launches else 0

    def _launch_substeps(self, actions_ptr, with_physics):
        """lr:393-434: action clip, then DEC x (torque -> simulate -> lag push)."""
        lib, p, b, r, st = self._lib, self._p_ref, self._b_ref, self._rng_ref(), self._stream()
        if with_physics:
            _lib.check(lib.ti5_begin_step(p, b, actions_ptr, st))
        for k in range(self._params.decimation):
            if with_physics:
                _lib.check(lib.ti5_torque_substep(p, b, r, k, st))
                self.gym.set_dof_actuation_force_tensor(self.sim, self.torques)
                self.gym.simulate(self.sim)
                self.gym.refresh_dof_state_tensor(self.sim)
                self.gym.refresh_actor_root_state_tensor(self.sim)
                _lib.check(lib.ti5_lag_push(p, b, k, st))
            else:
                # nothing runs between push k-1 and torque k: one fused launch per substep
                if k == 0:      # action clip fused into the first torque launch
                    _lib.check(lib.ti5_first_substep(p, b, r, actions_ptr, st))
                else:
                    _lib.check(lib.ti5_substep(p, b, r, k, C["TI5_SUB_TORQUE"] | C["TI5_SUB_PUSH"] | self._chain("TI5_SUB_CHAINED"), st))

    def _launch_post(self, with_physics, part=3, notify=False):
        """lr:458-506 post_physics_step + the observation clip of lr:441-446.  `part`: 1 = ti5_post_physics only,
        2 = ti5_reset_observe only (per-kernel timing), 3 = both.  `notify`: issue the simulator-facing calls of the
        phase where the reference issues them (not while capturing a graph: `step` issues them after the replay)."""
        lib, p, b, r, st = self._lib, self._p_ref, self._b_ref, self._rng_ref(), self._stream()
        fused = not with_physics
        if part & 1:
            if with_physics:
                self.gym.refresh_actor_root_state_tensor(self.sim)          # lr:464-466
                self.gym.refresh_net_contact_force_tensor(self.sim)
                self.gym.refresh_rigid_body_state_tensor(self.sim)
            if self._params.num_height_points:
                _lib.check(lib.ti5_sample_heights(p, b, st))
            _lib.check(lib.ti5_post_physics(p, b, r, (C["TI5_POST_PUSH_LAST"] | self._chain("TI5_POST_CHAINED")) if fused else 0, st))
            if notify:
                self._notify_simulator_of_disturbances()
        if part & 2:
            _lib.check(lib.ti5_reset_observe(p, b, r, C["TI5_RO_RESET"] | C["TI5_RO_OBSERVE"] | self._chain("TI5_RO_CHAINED"), st))
            if notify:
                self._notify_simulator_of_resets()

    def _materialize_windows(self):
        """lr:441-446 / t1:477-481: this step's windows as FRESH contiguous tensors (outside the captured graph: a
        caller may keep them across later steps, like the tensors `torch.cat` returns in the reference)."""
        p = self._params
        self._obs_out = torch.empty(self.num_envs, p.frame_stack * p.num_single_obs, dtype=torch.float32, device=self.device)
        self._priv_out = torch.empty(self.num_envs, p.c_frame_stack * p.priv_frame, dtype=torch.float32, device=self.device)
        b = _lib.Ti5Buffers.from_buffer_copy(self._buffers)
        b.obs_out, b.priv_out = self._obs_out.data_ptr(), self._priv_out.data_ptr()
        _lib.check(self._lib.ti5_materialize_obs(self._p_ref, ctypes.byref(b), self._stream()))

    def _launch_step(self, actions_ptr, with_physics, notify=False):
        """Enqueue the kernels of one policy step on the current stream."""
        if self._uses_fused_step(with_physics):
            # no simulator between the substeps: [heights] -> clip + DEC substeps + post-physics -> resets + observations
            lib, p, b, r, st = self._lib, self._p_ref, self._b_ref, self._rng_ref(), self._stream()
            opt = 0
            if self._params.num_height_points:
                _lib.check(lib.ti5_sample_heights(p, b, st))
                opt = self._chain("TI5_FUSED_CHAINED")
            _lib.check(lib.ti5_fused_step(p, b, r, actions_ptr, opt, st))
            if notify:
                self._notify_simulator_of_disturbances()
            self._launch_post(with_physics, part=2, notify=notify)
            return
        self._launch_substeps(actions_ptr, with_physics)
        self._launch_post(with_physics, notify=notify)

    def _uses_fused_step(self, with_physics=False):
        # (the per-substep re-draw of the action lag, off in t1_cfg, exists in the unfused substep kernels only)
        return (self._fused_step and not with_physics and self._params.env_block <= 64
                and not self._params.flags & C["TI5_F_LAG_PERSTEP"])

    def phase_launchers(self):
        """The kernel families of a fused step as [(name, launch, launches)], for a benchmark that brackets each family
        with CUDA events on the current stream (direct launches: the event records cut the programmatic chain anyway).
        Call all in order, then `_finish_step()`."""
        lib, p, b, r = self._lib, self._p_ref, self._b_ref, self._rng_ref()
        a_ptr = ctypes.c_void_p(self._actions_in.data_ptr())
        phases = []
        if self._params.num_height_points:
            phases.append(("heights", lambda: _lib.check(lib.ti5_sample_heights(p, b, self._stream())), 1))
        if self._uses_fused_step():
            phases.append(("fused_step", lambda: _lib.check(lib.ti5_fused_step(p, b, r, a_ptr, 0, self._stream())), 1))
        else:
            phases.append(("substep", lambda: self._launch_substeps(a_ptr, False), self._params.decimation))
            phases.append(("post_physics", lambda: _lib.check(lib.ti5_post_physics(
                p, b, r, C["TI5_POST_PUSH_LAST"] | self._chain("TI5_POST_CHAINED"), self._stream())), 1))
        phases.append(("reset_observe", lambda: self._launch_post(False, 2), 1))
        return phases

    def capture_phase_graphs(self):
        """One CUDA graph per kernel family of the step instead of one for the whole step, so that a benchmark can
        bracket each family with CUDA events: [(name, graph, launches)].  Replay all in order, then `_finish_step()`."""
        torch.cuda.synchronize(self.device)
        phases = self.phase_launchers()
        graphs = []
        for name, fn, n in phases:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=torch.cuda.Stream(self.device)):
                fn()
            graphs.append((name, g, n))
        return graphs

    @property
    def launches_per_step(self):
        step = 2 if self._uses_fused_step() else self._params.decimation + 2
        return step + (1 if self._params.num_height_points else 0) + (1 if self._materialize else 0)

    def step(self, actions):
        with_physics = getattr(self.gym, "physics", None) is not None or not hasattr(self.gym, "physics")
        if self._use_graph and not with_physics:
            # The captured step reads the actions where the caller left them: one graph per action buffer address (a
            # training loop hands over a handful of recurring allocations), so no copy kernel and no extra kernel
            # boundary in front of the step.  Other inputs (host tensors, odd dtypes / strides, too many distinct
            # addresses) are copied into a static buffer first.
            g = None
            if (actions.is_cuda and actions.dtype == torch.float32 and actions.is_contiguous()
                    and actions.device == self._actions_in.device and actions.shape == self._actions_in.shape):
                ptr = actions.data_ptr()
                g = self._graphs.get(ptr)
                if g is None and len(self._graphs) < self._max_graphs:
                    g = self._graphs[ptr] = self._capture_graph(ptr)
            if g is None:
                if self._graph is None:
                    self._graph = self._capture_graph(self._actions_in.data_ptr())
                self._actions_in.copy_(actions, non_blocking=True)     # device tensor or pinned host memory
                g = self._graph
            g.replay()
            # the simulator-facing calls of the step, in the reference's order; the kernels are already enqueued
            self._notify_simulator_of_disturbances()
            self._notify_simulator_of_resets()
        else:
            a = actions.to(device=self.device, dtype=torch.float32).contiguous()
            self._launch_step(ctypes.c_void_p(a.data_ptr()), with_physics, notify=True)
        return self._finish_step()

    # ------------------------------------------------------------------ host-side callers (CPU policy / controller)
    def enable_host_io(self):
        """Pinned host buffers wired into the captured step, for callers whose actions live in host memory (a CPU
        policy, a joystick loop like play.py:186-194).  The pinned pages are device-mapped (unified addressing), so
        the first substep kernel reads `host_actions` (N,12) f32 in place and ti5_reset_observe stores the packed
        per-step outputs [rew f32 | reset bool | time_outs bool] straight into `host_outputs`: no copy-engine hop, one
        graph launch per step.  Returns (host_actions, host_outputs)."""
        if not self._use_graph:
            raise _lib.Ti5Error("host I/O rides in the step's CUDA graph: needs rng_mode='philox', use_cuda_graph=True")
        N = self.num_envs
        self.host_actions = torch.zeros(N, self.num_actions, dtype=torch.float32).pin_memory()
        self.host_outputs = torch.zeros(6 * N, dtype=torch.uint8).pin_memory()
        self.host_rew = self.host_outputs[:4 * N].view(torch.float32)
        self.host_reset = self.host_outputs[4 * N:5 * N].view(torch.bool)
        self.host_time_outs = self.host_outputs[5 * N:].view(torch.bool)
        self._bind_buffers()
        self._drop_graphs()
        return self.host_actions, self.host_outputs

    def step_host(self):
        """`step()` for the actions currently in `host_actions`; returns once `host_rew` / `host_reset` /
        `host_time_outs` hold this step's results (a stream synchronise, the only host wait of the step)."""
        if self.host_actions is None:
            self.enable_host_io()
        if getattr(self.gym, "physics", None) is not None or not hasattr(self.gym, "physics"):
            raise _lib.Ti5Error("step_host() needs the graph-captured step (no simulator callbacks between substeps)")
        if self._graph_host is None:
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=torch.cuda.Stream(self.device)):
                self._launch_step(ctypes.c_void_p(self.host_actions.data_ptr()), False)
            self._graph_host = g
        self._graph_host.replay()
        self._notify_simulator_of_disturbances()
        self._notify_simulator_of_resets()
        out = self._finish_step()
        torch.cuda.current_stream(self.device).synchronize()
        return out

    def _drop_graphs(self):
        """Forget every captured step (buffers were re-bound or the launch sequence changed)."""
        self._graph = self._graph_host = None
        self._graphs = {}

    def _capture_graph(self, actions_ptr):
        torch.cuda.synchronize(self.device)
        side = torch.cuda.Stream(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self._launch_step(ctypes.c_void_p(actions_ptr), False)
        return g                # capture does not execute: device counters are untouched

    def _finish_step(self):
        """Host-side epilogue: counters, output views, extras (all without a device sync)."""
        self._step_index += 1
        self._view_epoch[0] = self._step_index
        self.common_step_counter += 1
        if hasattr(self.gym, "substep"):
            self.gym.substep = 0
        if self._materialize:
            self._materialize_windows()
            self.obs_buf, self.privileged_obs_buf = self._obs_out, self._priv_out
        else:
            self.obs_buf, self.privileged_obs_buf = self._window_views()
            # ring views die at the next step(): tag them, so that a storage that copies them late can tell
            # (`check_not_stale_ring_view`; the reference's runner holds the observation across a step, dh_ppo.py:88)
            self.obs_buf.ti5_ring_view = self.privileged_obs_buf.ti5_ring_view = (self._view_epoch, self._step_index)
        if self._params.log_len:
            # a fresh tensor object per step, tagged with the log row of its newest frame: the rollout storage
            # receives it one env step later (dh_ppo.py:88 -> rs:62) and must know which window it was
            tag = getattr(self.obs_buf, "ti5_ring_view", None)
            self.obs_buf = self.obs_buf.view_as(self.obs_buf)
            self.obs_buf.ti5_frame_row = self.frame_log_row
            if tag is not None:
                self.obs_buf.ti5_ring_view = tag
        self._publish_extras()
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def _publish_extras(self):
        names, trimesh, curriculum, timeouts = self._extras_args
        self.extras["episode"] = EpisodeInfo(self._extras_rows[self._step_index % len(self._extras_rows)], names, trimesh, curriculum)
        if timeouts:
            self.extras["time_outs"] = self._time_outs_latched

    # ------------------------------------------------------------------ pieces, with the reference's names
    def _compute_torques(self, actions, substep=0):
        """lr:1019-1074 for one substep (advanced use; `step` drives the fused sequence)."""
        a = actions.to(device=self.device, dtype=torch.float32).contiguous()
        self.actions.copy_(a)
        _lib.check(self._lib.ti5_torque_substep(self._p_ref, self._b_ref, self._rng_ref(), substep, self._stream()))
        return self.torques

    def _get_heights(self, env_ids=None):
        """lr:1551-1587."""
        _lib.check(self._lib.ti5_sample_heights(self._p_ref, self._b_ref, self._stream()))
        return self.measured_heights

    def post_physics_step(self):
        """lr:458-506 on the current simulator state: termination, rewards, resets, observations."""
        with_physics = getattr(self.gym, "physics", None) is not None or not hasattr(self.gym, "physics")
        self._launch_post(with_physics)

    def compute_observations(self):
        """t1:368-481 on the current state (no resets)."""
        _lib.check(self._lib.ti5_reset_observe(self._p_ref, self._b_ref, self._rng_ref(), C["TI5_RO_OBSERVE"], self._stream()))

    def reset_idx(self, env_ids):
        """lr:520-602 / t1:483-559 for an explicit id list (the in-step resets are fused into `step`)."""
        if len(env_ids) == 0:
            return
        self.reset_buf.zero_()
        self.reset_buf[env_ids.long()] = True
        st = self._stream()
        _lib.check(self._lib.ti5_reset_bookkeeping(self._p_ref, self._b_ref, st))
        _lib.check(self._lib.ti5_reset_scatter(self._p_ref, self._b_ref, self._rng_ref(), st))
        self._notify_simulator_of_resets()
        self._publish_extras()

    # ------------------------------------------------------------------ lower boundary: what the simulator is told
    def _notify_simulator_of_disturbances(self):
        """Hook of the task class: hand the disturbances of `_post_physics_step_callback` to the simulator (the base
        callback of the reference, lr:520-560 region, has none that touch the simulator in t1's configuration)."""

    def _notify_simulator_of_resets(self):
        """lr:1087-1090, 1117-1120, 915-939 for the envs ti5_reset_observe re-spawned: joint state, root state and the
        re-drawn joint properties go to the simulator, addressed by the ascending id list the kernel left in
        `reset_ids`.  A binding that declares `device_counts` gets the list as (buffer, device count) and the step
        stays free of host waits; Isaac Gym's own API needs `len(env_ids)` on the host — one 4-byte read-back, where
        the reference pays `nonzero()` (lr:490)."""
        gym = self.gym
        if getattr(gym, "device_counts", False):
            n = self.n_reset_device
            gym.set_dof_state_tensor_indexed(self.sim, self.dof_state, self.reset_ids, n)
            gym.set_actor_root_state_tensor_indexed(self.sim, self.root_states, self.reset_ids, n)
            self._refresh_actor_dof_props(None, n)
            return
        if not getattr(gym, "needs_indexed_resets", True):
            return
        n = int(self.n_reset_device)                        # the step's only host wait with a real simulator
        if n:
            ids = self.reset_ids[:n]
            gym.set_dof_state_tensor_indexed(self.sim, self.dof_state, ids, n)
            gym.set_actor_root_state_tensor_indexed(self.sim, self.root_states, ids, n)
            self._refresh_actor_dof_props(None, n)
            gym.refresh_actor_root_state_tensor(self.sim)       # t1:544-546
            gym.refresh_net_contact_force_tensor(self.sim)
            gym.refresh_rigid_body_state_tensor(self.sim)

    def _refresh_actor_dof_props(self, env_ids=None, n=None):
        """lr:915-939.  The reference walks `env_ids` in Python and reads `joint_armatures[env_id, i]` element by
        element (a device->host sync each with a GPU pipeline).  Here the properties of the listed envs are one dense
        (n, 12, 3) tensor [friction multiplier, damping multiplier, armature]: written by the reset scatter for the
        envs of `reset_ids` (env_ids=None), or gathered for any id list by ti5_gather_dof_props.  A simulator binding
        with `set_actor_dof_properties_batched` takes the tensor as it is; Isaac Gym's per-env property structs are
        filled from ONE host copy of it."""
        gym, dr = self.gym, self.cfg.domain_rand
        if not (dr.randomize_joint_armature or getattr(dr, "randomize_joint_friction", False)
                or getattr(dr, "randomize_joint_damping", False)):
            return
        ids, props = self.reset_ids, self._dof_props
        if env_ids is not None:
            ids = torch.as_tensor(env_ids, device=self.device).to(torch.int32).contiguous()
            n = len(ids)
            if n == 0:
                return
            count = torch.tensor([n], dtype=torch.int32, device=self.device)
            props = torch.empty(n, self.num_dof, 3, dtype=torch.float32, device=self.device)
            _lib.check(self._lib.ti5_gather_dof_props(self._p_ref, self._b_ref, ctypes.c_void_p(ids.data_ptr()),
                                                      ctypes.c_void_p(count.data_ptr()), n,
                                                      ctypes.c_void_p(props.data_ptr()), self._stream()))
        batched = getattr(gym, "set_actor_dof_properties_batched", None)
        if batched is not None:
            batched(self.sim, ids, props, n)
            return
        if not hasattr(gym, "set_actor_dof_properties"):
            return
        n = int(n)
        ids_h, props_h = ids[:n].cpu().tolist(), props[:n].cpu().numpy()
        for r, env_id in enumerate(ids_h):
            dp = gym.get_actor_dof_properties(self.envs[env_id], 0)
            if getattr(dr, "randomize_joint_friction", False):
                dp["friction"] *= props_h[r, :, 0]
            if getattr(dr, "randomize_joint_damping", False):
                dp["damping"] *= props_h[r, :, 1]
            if dr.randomize_joint_armature:
                dp["armature"][:] = props_h[r, :, 2]
            gym.set_actor_dof_properties(self.envs[env_id], 0, dp)

    # ------------------------------------------------------------------ curriculum state mirrored from the device
    def sync_from_device(self):
        """Refresh the host mirrors of device-resident scalars (command ranges, reset count)."""
        g = self._read_globals()
        for i, key in enumerate(("lin_vel_x", "lin_vel_y", "ang_vel_yaw")):
            cur = g.cmd_range[(g.step_index + 1) & 1][i]      # the copy the next step will read
            self.command_ranges[key] = [cur[0], cur[1]]
        self.num_resets_last_step = int(g.n_reset)
        return g

    def set_common_step_counter(self, value):
        self.common_step_counter = int(value)
        g = self._read_globals()
        g.common_step_offset = self.common_step_counter - self._step_index
        self._globals.copy_(torch.frombuffer(bytearray(bytes(g)), dtype=torch.uint8))

    # ------------------------------------------------------------------ state import (tests, resume)
    def load_state(self, state):
        """Adopt a full per-env state given with the reference's attribute names and layouts
        (lag buffers as shifted (N, W, len) arrays, histories oldest -> newest)."""
        N = self.num_envs
        dev = self.device
        plain = ("torques actions last_actions last_last_actions last_dof_vel last_root_vel commands feet_air_time "
                 "feet_height last_feet_z base_quat base_lin_vel base_ang_vel projected_gravity base_euler_xyz "
                 "feet_euler_xyz ext_forces ext_torques rand_push_force rand_push_torque ref_dof_pos ref_action gait_time gait_start "
                 "torque_multi motor_offsets randomized_p_gains randomized_d_gains randomized_joint_coulomb "
                 "randomized_joint_viscous joint_armatures phase_length_buf rew_buf env_origins env_frictions body_mass "
                 "last_contacts contact_filt time_out_buf").split()
        for name in plain:
            if name in state:
                getattr(self, name).copy_(torch.as_tensor(state[name]).to(dev).view_as(getattr(self, name)))
        self._episode_length_buf.copy_(torch.as_tensor(state["episode_length_buf"]).to(dev))
        self.reset_buf.copy_(torch.as_tensor(state["reset_buf"]).to(dev).bool())
        for col, name in enumerate(("lag_timestep", "dof_lag_timestep", "imu_lag_timestep")):
            self._lag_timestep[:, col] = torch.as_tensor(state[name]).to(dev).to(torch.int32)
        lag_opts_on = self._last_lag.shape[1] == N and N > 1 or self._last_lag.shape[1] == N == 1
        if self._params.flags2:
            for col, name in enumerate(("joint_friction_coeffs", "joint_damping_coeffs")):
                if name in state:
                    self._joint_coeffs[:, col] = torch.as_tensor(state[name]).to(dev).float().view(-1)
        for col, name in enumerate(("dof_pos_lag_timestep", "dof_vel_lag_timestep")):
            if name in state and lag_opts_on:
                self._lag_pv[:, col] = torch.as_tensor(state[name]).to(dev).to(torch.int32)
        for col, name in enumerate(("last_lag_timestep", "last_dof_lag_timestep", "last_imu_lag_timestep",
                                    "last_dof_pos_lag_timestep", "last_dof_vel_lag_timestep")):
            if name in state and lag_opts_on:
                self._last_lag[:, :, col] = torch.as_tensor(state[name]).to(dev).to(torch.int32)
        for i, name in enumerate(self.reward_scales):
            self.episode_sums[name].copy_(torch.as_tensor(state["episode_sums"][i]).to(dev))
        if "terrain_levels" in state:
            self.terrain_levels.copy_(torch.as_tensor(state["terrain_levels"]).to(dev))
            self.terrain_types.copy_(torch.as_tensor(state["terrain_types"]).to(dev))
        if "terrain_origins" in state:
            self.terrain_origins.copy_(torch.as_tensor(state["terrain_origins"]).to(dev).float())
        cnt = [int(v) for v in state["counters"]]
        self.common_step_counter = cnt[0]
        cr = torch.as_tensor(state["command_ranges"]).tolist()
        for i, key in enumerate(("lin_vel_x", "lin_vel_y", "ang_vel_yaw")):
            self.command_ranges[key] = list(cr[i])
        self._step_index = STEP_INDEX0
        self._write_globals()
        g = self._read_globals()
        g.is_first_add_force[0] = g.is_first_add_force[1] = cnt[1]
        self._globals.copy_(torch.frombuffer(bytearray(bytes(g)), dtype=torch.uint8))
        # lag buffers: age a (slot 0 = newest) was push index pushes-1-a
        pushes = self._step_index * self._params.decimation
        self._ring_stamp.zero_()
        pos_vel = bool(self._params.flags & C["TI5_F_POS_VEL_LAG"])
        for ring, name in ((self._act_ring, "lag_buffer"), (self._dof_ring, "dof_lag_buffer"), (self._imu_ring, "imu_lag_buffer")):
            if pos_vel and name == "dof_lag_buffer":                # the DOF ring holds the position and the velocity lag buffers
                pb, vb = (torch.as_tensor(state[k]).to(dev) for k in ("dof_pos_lag_buffer", "dof_vel_lag_buffer"))
                L = ring.shape[0]
                pad = lambda t: torch.nn.functional.pad(t, (0, L - t.shape[2]))
                buf = torch.cat((pad(pb), pad(vb)), 1)
            else:
                buf = torch.as_tensor(state[name]).to(dev)          # (N, W, len)
            length = buf.shape[2]
            assert pushes >= length
            slots = (pushes - 1 - torch.arange(length, device=dev)) % length
            ring[slots] = buf.permute(2, 0, 1)
        # histories: frame i (0 = oldest) sits at slot (step_index + i) % H of the mirrored ring
        for ring, name in ((self._obs_ring, "obs_history"), (self._priv_ring, "critic_history")):
            hist = torch.as_tensor(state[name]).to(dev)             # (H, N, K)
            Hn = hist.shape[0]
            slots = (self._step_index + torch.arange(Hn, device=dev)) % Hn
            lim = self.cfg.normalization.clip_observations
            fr = torch.clip(hist, -lim, lim).permute(1, 0, 2)       # the ring stores clipped frames
            ring[:, slots] = fr
            ring[:, slots + Hn] = fr
        self._drop_graphs()
        if self._params.log_len:
            self._seed_frame_log()
        self.obs_buf, self.privileged_obs_buf = self._history_views()
