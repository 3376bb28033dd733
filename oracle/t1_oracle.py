"""CPU/torch ORACLE for the `t1_dh_stand` step math and the GAE scan.

TEST INFRASTRUCTURE — not product code.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; the shipped
package `ti5_isaacgym_b200` never does (its CUDA extension has no CPU fallback).

What it is: a restatement, in plain batched torch ops, of the algorithm the reference
implements in
    humanoid/envs/base/legged_robot.py      ("lr")   step / torques / termination / reset / heights
    humanoid/envs/t1/t1_dh_stand_env.py     ("t1")   gait phase, rewards, observations, T1 reset
    humanoid/utils/math.py                           quat_apply_yaw
    humanoid/algo/ppo/rollout_storage.py             compute_returns (GAE)
plus the six `isaacgym.torch_utils` helpers (Isaac Gym Preview 4: third party, un-vendored,
unpinned — reference setup.py:11), restated from their published definitions.  Each function
cites the reference lines it follows.

Parity status: PINNED against the reference itself.  The reference has no tests or golden
vectors, so `oracle/pin_against_reference.py` imports the unmodified
reference from /root/reference (under `oracle/shim`), drives both with identical synthetic
simulator tensors and identical random draws and requires bit-equal results on CPU; the
same script writes the fixtures in `tests/golden/`.

Randomness is an INPUT ("RNG-as-input", SURVEY section 7 hard part 1): every draw site
consumes a per-env uniform from the pool dict `R` (see `rng_pool_shapes`); an env that does
not draw at a site ignores its entry.  Layout conventions follow the reference (lag buffers
are shifted arrays with slot 0 = newest, histories oldest -> newest).
"""
from types import SimpleNamespace

import numpy as np
import torch

# ------------------------------------------------------------------------------------------
# isaacgym.torch_utils restatements (published Preview-4 definitions)
# ------------------------------------------------------------------------------------------


def quat_rotate_inverse(q, v):
    qw = q[:, -1]
    qv = q[:, :3]
    a = v * (2.0 * qw ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(qv, v, dim=-1) * qw.unsqueeze(-1) * 2.0
    c = qv * torch.bmm(qv.view(-1, 1, 3), v.view(-1, 3, 1)).squeeze(-1) * 2.0
    return a - b + c


def quat_apply(q, v):
    shape = v.shape
    q = q.reshape(-1, 4)
    v = v.reshape(-1, 3)
    xyz = q[:, :3]
    t = xyz.cross(v, dim=-1) * 2
    return (v + q[:, 3:] * t + xyz.cross(t, dim=-1)).view(shape)


def quat_apply_yaw(q, v):
    """humanoid/utils/math.py:8-12 (normalize = x / norm.clamp(min=1e-9))."""
    qy = q.clone().view(-1, 4)
    qy[:, :2] = 0.
    qy = qy / qy.norm(p=2, dim=-1).clamp(min=1e-9, max=None).unsqueeze(-1)
    return quat_apply(qy, v)


def euler_xyz(q):
    """lr:27-53 / t1:16-39: XYZ Euler angles of xyzw quaternions, wrapped to (-pi, pi]."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    roll = torch.atan2(2.0 * (w * x + y * z), w * w - x * x - y * y + z * z)
    sinp = 2.0 * (w * y - z * x)
    half_pi = torch.tensor(np.pi / 2.0, device=q.device, dtype=torch.float).expand_as(sinp)
    pitch = torch.where(torch.abs(sinp) >= 1, torch.abs(half_pi) * torch.sign(sinp), torch.asin(sinp))
    yaw = torch.atan2(2.0 * (w * z + x * y), w * w + x * x - y * y - z * z)
    e = torch.stack((roll % (2 * np.pi), pitch % (2 * np.pi), yaw % (2 * np.pi)), dim=-1)
    e[e > np.pi] -= 2 * np.pi
    return e


# ------------------------------------------------------------------------------------------
# constants
# ------------------------------------------------------------------------------------------

GAIT_KINDS = ("stand", "walk_sagittal", "walk_lateral", "rotate", "walk_omnidirectional")
DR_ROWS = ("torque_multi", "motor_offset", "p_gain", "d_gain", "coulomb", "viscous", "armature")


def public_dict(obj):
    """`class_to_dict` (humanoid/utils/helpers.py:14-29): `dir()` order = alphabetical."""
    if not hasattr(obj, "__dict__"):
        return obj
    return {k: ([public_dict(i) for i in getattr(obj, k)] if isinstance(getattr(obj, k), list)
                else public_dict(getattr(obj, k)))
            for k in dir(obj) if not k.startswith("_")}


def make_consts(cfg, sim_dt, robot, device="cpu", terrain=None):
    """Everything `_parse_cfg` (lr:94-113), `_init_buffers` (lr:212-249),
    `_prepare_reward_function` (lr:352-384) and `_process_dof_props` (lr:837-849) derive."""
    C = SimpleNamespace(cfg=cfg, device=device, robot=robot, terrain=terrain)
    f32 = dict(dtype=torch.float, device=device)
    C.dt = cfg.control.decimation * sim_dt
    C.decimation = cfg.control.decimation
    C.max_episode_length_s = cfg.env.episode_length_s
    C.max_episode_length = np.ceil(C.max_episode_length_s / C.dt)
    C.push_interval = np.ceil(cfg.domain_rand.push_interval_s / C.dt)
    C.ext_force_interval = np.ceil(cfg.domain_rand.ext_force_interval_s / C.dt)
    C.obs_scales = cfg.normalization.obs_scales
    scales = public_dict(cfg.rewards.scales)
    C.reward_scales = {k: v * C.dt for k, v in scales.items() if v != 0}     # lr:357-364
    C.reward_names = [k for k in C.reward_scales if k != "termination"]
    C.num_dof = robot.num_dof
    C.num_bodies = robot.num_bodies
    C.feet = torch.tensor(robot.feet_indices, dtype=torch.long, device=device)
    C.knees = torch.tensor(robot.knee_indices, dtype=torch.long, device=device)
    C.pen_bodies = torch.tensor(robot.penalised_contact_indices, dtype=torch.long, device=device)
    C.term_bodies = torch.tensor(robot.termination_contact_indices, dtype=torch.long, device=device)
    # lr:216-234 (default pose, substring-matched PD gains), lr:843-849 (limits)
    C.default_dof_pos = torch.zeros(1, C.num_dof, **f32)
    C.p_gains = torch.zeros(C.num_dof, **f32)
    C.d_gains = torch.zeros(C.num_dof, **f32)
    C.torque_limits = torch.zeros(C.num_dof, **f32)
    C.dof_vel_limits = torch.zeros(C.num_dof, **f32)
    for i, name in enumerate(robot.dof_names):
        C.default_dof_pos[0, i] = cfg.init_state.default_joint_angles[name]
        for key in cfg.control.stiffness:
            if key in name:
                C.p_gains[i] = cfg.control.stiffness[key]
                C.d_gains[i] = cfg.control.damping[key]
        C.torque_limits[i] = float(np.float32(robot.dof_effort[i])) * cfg.safety.torque_limit
        C.dof_vel_limits[i] = float(np.float32(robot.dof_velocity[i])) * cfg.safety.vel_limit
    C.commands_scale = torch.tensor([C.obs_scales.lin_vel, C.obs_scales.lin_vel, C.obs_scales.ang_vel], device=device)
    C.command_ranges = public_dict(cfg.commands.ranges)
    st = cfg.init_state
    C.base_init_state = torch.tensor(st.pos + st.rot + st.lin_vel + st.ang_vel, **f32)      # lr:1338
    C.gravity_vec = torch.tensor([0., 0., -1.], **f32)                                        # lr:168
    # noise vector, t1:326-357
    K, D, nc = cfg.env.num_single_obs, cfg.env.num_actions, cfg.env.num_commands
    ns_, os_ = cfg.noise.noise_scales, C.obs_scales
    nv = torch.zeros(K, device=device)
    nv[nc:nc + D] = ns_.dof_pos * os_.dof_pos
    nv[nc + D:nc + 2 * D] = ns_.dof_vel * os_.dof_vel
    nv[nc + 3 * D:nc + 3 * D + 3] = ns_.ang_vel * os_.ang_vel
    nv[nc + 3 * D + 3:nc + 3 * D + 6] = ns_.quat * os_.quat
    C.noise_scale_vec = nv
    dr = cfg.domain_rand
    # options t1_cfg marks "always False" (t1_cfg:290-312), restated all the same: per-step re-draws of the lag indices
    # (lr:1038-1043, t1:408-413, 417-430, 437-442) and separate position / velocity lags (lr:425-430, t1:416-431)
    C.perstep = {k: bool(getattr(dr, f"randomize_{k}_timesteps_perstep", False) and getattr(dr, f"randomize_{k}_timesteps", False))
                 for k in ("lag", "dof_lag", "imu_lag", "dof_pos_lag", "dof_vel_lag")}
    C.pos_vel_lag = bool(getattr(dr, "add_dof_pos_vel_lag", False))
    # joint friction / damping multipliers (lr:755-773, 921-931): one per env; the `_each_joint` variants read per-joint
    # ranges t1_cfg only has for ten of the twelve joints (the reference raises AttributeError at joint 11)
    C.joint_props = (bool(getattr(dr, "randomize_joint_friction", False)), bool(getattr(dr, "randomize_joint_damping", False)))
    assert not (C.joint_props[0] and dr.randomize_joint_friction_each_joint) and \
        not (C.joint_props[1] and dr.randomize_joint_damping_each_joint), "per-joint ranges exist for ten joints only"
    C.heading_command = bool(cfg.commands.heading_command)
    C.forward_vec = torch.tensor([1., 0., 0.], **f32)                                         # lr:170
    C.custom_origins = cfg.terrain.mesh_type in ("heightfield", "trimesh")                   # lr:1481
    C.curriculum = cfg.terrain.curriculum and C.custom_origins                                 # lr:104-105
    if cfg.terrain.measure_heights:                                                            # lr:1535-1549
        y = torch.tensor(cfg.terrain.measured_points_y, device=device)
        x = torch.tensor(cfg.terrain.measured_points_x, device=device)
        gx, gy = torch.meshgrid(x, y, indexing="ij")
        C.height_points = torch.stack((gx.flatten(), gy.flatten(), torch.zeros(gx.numel(), device=device)), dim=1)
    return C


def rng_pool_shapes(C, N):
    """Per-step uniform / integer pools (name -> (shape, kind)).  kind 'u' = U[0,1) fp32,
    otherwise ('i', low, high) = integers in [low, high)."""
    D, K = C.num_dof, C.cfg.env.num_single_obs
    dr = C.cfg.domain_rand
    G = len(C.cfg.commands.gait)                      # gait slots of the schedule (three in t1_cfg)
    return {
        "torque": ((C.decimation, N, D), "u"),        # lr:1071, redrawn every substep (A18)
        "cmd": ((2, G, N, 3), "u"),                   # t1:126-177; [0]=callback pass, [1]=pass inside reset_idx
        "push": ((N, 5), "u"),                        # t1:223-226
        "ext": ((N, 6), "u"),                         # t1:237-241
        "dofs": ((N, D), "u"),                        # lr:1084
        "root_xy": ((N, 2), "u"),                     # lr:1105-1108
        "dr": ((N, len(DR_ROWS), D), "u"),            # lr:735-783
        "gait_time": ((N, G), "u"),                   # t1:116
        "noise": ((N, K), "u"),                       # t1:472
        "lag_idx": ((N, 3), ("i", 0, 0)),             # lr:608-629 (ranges applied per column by the caller)
        "gait_start": ((N,), ("i", 0, 2)),            # t1:523 (CPU generator in the reference, A25)
        "terrain_level": ((N,), ("i", 0, max(1, getattr(C.cfg.terrain, "num_rows", 1)))),  # lr:1156
    } | ({"dr_joint": ((N, 2), "u")} if any(C.joint_props) else {}) \
      | ({"lag_idx_pv": ((N, 2), ("i", 0, 0))} if C.pos_vel_lag else {}) \
      | ({"lag_step": ((C.decimation + 4, N), ("i", 0, 0))} if any(C.perstep.values()) else {})
    # lag_idx_pv: lr:639, 646 (position / velocity lag at a reset); lag_step: per-step re-draws — rows 0 .. DEC-1 the
    # action lag of each substep (lr:1039), then the DOF, IMU, position and velocity lags of the step (t1:409, 438, 418, 426)


def draw_pools(C, N, gen, device="cpu"):
    """Seeded pools for one step (CPU generator, then moved)."""
    dr = C.cfg.domain_rand
    out = {}
    for name, (shape, kind) in rng_pool_shapes(C, N).items():
        if kind == "u":
            out[name] = torch.rand(*shape, generator=gen).to(device)
        elif name == "lag_idx":
            cols = [torch.randint(r[0], r[1] + 1, (N,), generator=gen) for r in
                    (dr.lag_timesteps_range, dr.dof_lag_timesteps_range, dr.imu_lag_timesteps_range)]
            out[name] = torch.stack(cols, 1).to(device)
        elif name == "lag_idx_pv":
            cols = [torch.randint(r[0], r[1] + 1, (N,), generator=gen) for r in
                    (dr.dof_pos_lag_timesteps_range, dr.dof_vel_lag_timesteps_range)]
            out[name] = torch.stack(cols, 1).to(device)
        elif name == "lag_step":
            rows = [dr.lag_timesteps_range] * C.decimation + [dr.dof_lag_timesteps_range, dr.imu_lag_timesteps_range,
                                                              dr.dof_pos_lag_timesteps_range, dr.dof_vel_lag_timesteps_range]
            out[name] = torch.stack([torch.randint(r[0], r[1] + 1, (N,), generator=gen) for r in rows], 0).to(device)
        else:
            out[name] = torch.randint(kind[1], kind[2], shape, generator=gen).to(device)
    return out


def _affine(lo, hi, u):
    """`torch_rand_float(lo, hi, ...)` = (hi - lo) * rand + lo with Python-double bounds."""
    return (hi - lo) * u + lo


# ------------------------------------------------------------------------------------------
# state
# ------------------------------------------------------------------------------------------


def new_state(C, N):
    """Zero/default-initialised buffers: base_task.py:55-74, lr:160-349, lr:1420-1474,
    t1:75-77, 562-569.  Random construction-time draws are NOT made here: tests copy them
    from the reference env (`adopt_reference_state`) or use `randomize_initial_state`."""
    cfg, dev = C.cfg, C.device
    D, NB = C.num_dof, C.num_bodies
    dr = cfg.domain_rand
    z = lambda *s, **k: torch.zeros(*s, device=dev, **({"dtype": torch.float} | k))
    S = SimpleNamespace(N=N)
    S.common_step_counter = 0
    S.is_first_add_force = True
    S.is_first_push = True
    S.init_done = True
    S.command_ranges = {k: list(v) for k, v in C.command_ranges.items()}
    S.rew_buf = z(N)
    S.reset_buf = torch.ones(N, device=dev, dtype=torch.long)
    S.time_out_buf = torch.zeros(N, device=dev, dtype=torch.bool)
    S.episode_length_buf = torch.zeros(N, device=dev, dtype=torch.long)
    S.phase_length_buf = torch.zeros(N, device=dev, dtype=torch.long)
    S.extras = {}
    S.torques, S.actions, S.last_actions, S.last_last_actions = z(N, D), z(N, D), z(N, D), z(N, D)
    S.last_dof_vel, S.last_root_vel = z(N, D), z(N, 6)
    S.commands = z(N, cfg.commands.num_commands)
    S.feet_air_time, S.feet_height = z(N, 2), z(N, 2)
    S.last_feet_z = 0                                   # Python int until the first step (A9)
    S.last_contacts = torch.zeros(N, 2, dtype=torch.bool, device=dev)
    S.contact_filt = torch.zeros(N, 2, dtype=torch.bool, device=dev)
    S.base_quat = z(N, 4)
    S.base_quat[:, 3] = 1
    S.base_lin_vel, S.base_ang_vel, S.projected_gravity = z(N, 3), z(N, 3), z(N, 3)
    S.projected_gravity[:, 2] = -1
    S.base_euler_xyz, S.feet_euler_xyz = z(N, 3), z(N, 2, 3)
    S.rand_push_force, S.rand_push_torque, S.ext_forces, S.ext_torques = z(N, 3), z(N, 3), z(N, 3), z(N, 3)
    S.applied_force, S.applied_torque = z(N, 3), z(N, 3)   # body-0 rows handed to apply_rigid_body_force_tensors (t1:247)
    S.ref_dof_pos, S.ref_action = z(N, D), z(N, D)
    S.measured_heights = 0
    S.gait_time = torch.zeros(N, len(cfg.commands.gait), dtype=torch.int, device=dev)
    S.gait_start = z(N)
    S.env_frictions, S.body_mass = z(N, 1), z(N, 1)
    S.env_origins = z(N, 3)
    S.torque_multi = torch.ones(N, D, device=dev)
    S.motor_offsets, S.randomized_p_gains, S.randomized_d_gains = z(N, D), z(N, D), z(N, D)
    S.randomized_joint_coulomb, S.randomized_joint_viscous, S.joint_armatures = z(N, D), z(N, D), z(N, D)
    S.joint_friction_coeffs, S.joint_damping_coeffs = torch.ones(N, 1, device=dev), torch.ones(N, 1, device=dev)   # lr:1449-1457
    S.lag_buffer = z(N, D, dr.lag_timesteps_range[1] + 1)
    S.dof_lag_buffer = z(N, 2 * D, dr.dof_lag_timesteps_range[1] + 1)
    S.imu_lag_buffer = z(N, 6, dr.imu_lag_timesteps_range[1] + 1)
    S.lag_timestep = torch.full((N,), dr.lag_timesteps_range[1], dtype=torch.long, device=dev)
    S.dof_lag_timestep = torch.full((N,), dr.dof_lag_timesteps_range[1], dtype=torch.long, device=dev)
    S.imu_lag_timestep = torch.full((N,), dr.imu_lag_timesteps_range[1], dtype=torch.long, device=dev)
    S.dof_pos_lag_buffer = z(N, D, dr.dof_pos_lag_timesteps_range[1] + 1)
    S.dof_vel_lag_buffer = z(N, D, dr.dof_vel_lag_timesteps_range[1] + 1)
    S.dof_pos_lag_timestep = torch.full((N,), dr.dof_pos_lag_timesteps_range[1], dtype=torch.long, device=dev)
    S.dof_vel_lag_timestep = torch.full((N,), dr.dof_vel_lag_timesteps_range[1], dtype=torch.long, device=dev)
    for k in ("lag", "dof_lag", "imu_lag", "dof_pos_lag", "dof_vel_lag"):       # lr:282, 301, 317, 335, 345
        setattr(S, f"last_{k}_timestep", torch.full((N,), getattr(dr, f"{k}_timesteps_range")[1], dtype=torch.long, device=dev))
    H, CH, K = cfg.env.frame_stack, cfg.env.c_frame_stack, cfg.env.num_single_obs
    P = cfg.env.single_num_privileged_obs + (cfg.terrain.num_height if cfg.terrain.measure_heights else 0)
    S.obs_history = z(H, N, K)          # oldest -> newest (the reference's deque, lr:251-267)
    S.critic_history = z(CH, N, P)
    S.obs_buf, S.privileged_obs_buf = z(N, H * K), z(N, CH * P)
    S.episode_sums = {k: z(N) for k in C.reward_scales}
    if C.custom_origins:
        S.terrain_levels = torch.zeros(N, dtype=torch.long, device=dev)
        S.terrain_types = torch.zeros(N, dtype=torch.long, device=dev)
    return S


def sim_views(sim, N, D, NB):
    """The wrapper views of lr:141-154 over the AoS gym tensors."""
    dof = sim.dof_state.view(N, D, 2)
    return dof[..., 0], dof[..., 1], sim.contact_forces.view(N, NB, 3), sim.rigid_state.view(N, NB, 13)


# ------------------------------------------------------------------------------------------
# substep phase (lr:399-434)
# ------------------------------------------------------------------------------------------


def _redraw_lag(S, kind, draw):
    """The per-step re-draw of a lag index (lr:1039-1043 and its four copies in t1:408-442): a fresh draw, but never more
    than one step further back than the last one."""
    new = draw.clone()
    last = getattr(S, f"last_{kind}_timestep")
    cond = new > last + 1
    new[cond] = last[cond] + 1
    setattr(S, f"{kind}_timestep", new)
    setattr(S, f"last_{kind}_timestep", new.clone())


def torque_substep(C, S, sim, actions, u_torque, lag_step=None):
    """lr:1019-1074: lagged PD with motor offset, viscous + Coulomb friction, fresh motor
    strength multiplier, clip to the torque limits."""
    dr = C.cfg.domain_rand
    q, qd, _, _ = sim_views(sim, S.N, C.num_dof, C.num_bodies)
    target = actions * C.cfg.control.action_scale
    if dr.add_lag:
        S.lag_buffer[:, :, 1:] = S.lag_buffer[:, :, :dr.lag_timesteps_range[1]].clone()
        S.lag_buffer[:, :, 0] = target.clone()
        if C.perstep["lag"]:                                               # lr:1038-1043
            _redraw_lag(S, "lag", lag_step)
        target = S.lag_buffer[torch.arange(S.N, device=C.device), :, S.lag_timestep.int()]
    S.lagged_actions_scaled = target
    kp, kd = (S.randomized_p_gains, S.randomized_d_gains) if dr.randomize_gains else (C.p_gains, C.d_gains)
    tau = kp * (target + C.default_dof_pos - q + S.motor_offsets) - kd * qd
    if dr.randomize_coulomb_friction:
        tau = tau - S.randomized_joint_viscous * qd - S.randomized_joint_coulomb * torch.sign(qd)
    if dr.randomize_torque:
        lo, hi = dr.torque_multiplier_range
        S.torque_multi = _affine(lo, hi, u_torque)
        tau = tau * S.torque_multi
    return torch.clip(tau, -C.torque_limits, C.torque_limits)


def lag_push(C, S, sim):
    """lr:412-434: push (q, qd) and (body-frame angular velocity, Euler angles) into the
    DOF / IMU lag buffers after every simulator substep."""
    dr = C.cfg.domain_rand
    q, qd, _, _ = sim_views(sim, S.N, C.num_dof, C.num_bodies)
    if dr.add_dof_lag:
        S.dof_lag_buffer[:, :, 1:] = S.dof_lag_buffer[:, :, :dr.dof_lag_timesteps_range[1]].clone()
        S.dof_lag_buffer[:, :, 0] = torch.cat((q, qd), 1)
    if C.pos_vel_lag:                                                      # lr:425-430
        S.dof_pos_lag_buffer[:, :, 1:] = S.dof_pos_lag_buffer[:, :, :dr.dof_pos_lag_timesteps_range[1]].clone()
        S.dof_pos_lag_buffer[:, :, 0] = q.clone()
        S.dof_vel_lag_buffer[:, :, 1:] = S.dof_vel_lag_buffer[:, :, :dr.dof_vel_lag_timesteps_range[1]].clone()
        S.dof_vel_lag_buffer[:, :, 0] = qd.clone()
    if dr.add_imu_lag:
        S.base_quat[:] = sim.root_states[:, 3:7]
        S.base_ang_vel[:] = quat_rotate_inverse(S.base_quat, sim.root_states[:, 10:13])
        S.base_euler_xyz = euler_xyz(S.base_quat)
        S.imu_lag_buffer[:, :, 1:] = S.imu_lag_buffer[:, :, :dr.imu_lag_timesteps_range[1]].clone()
        S.imu_lag_buffer[:, :, 0] = torch.cat((S.base_ang_vel, S.base_euler_xyz), 1)


# ------------------------------------------------------------------------------------------
# gait phase helpers (t1:80-107, 250-274)
# ------------------------------------------------------------------------------------------


def stand_command(C, S):
    return torch.norm(S.commands[:, :3], dim=1) <= C.cfg.commands.stand_com_threshold


def gait_phase(C, S):
    """t1:80-92.  Side effect kept: standing envs get `phase_length_buf = 0` (A4)."""
    if not C.cfg.commands.sw_switch:                                           # t1:89-90
        return (S.episode_length_buf * C.dt / C.cfg.rewards.cycle_time) % 1.0 + S.gait_start
    stand = stand_command(C, S)
    S.phase_length_buf[stand] = 0
    return ((S.phase_length_buf * C.dt / C.cfg.rewards.cycle_time) % 1.0 + S.gait_start) * (~stand)


def stance_mask(C, S):
    """t1:95-107: 1 = stance, 0 = swing; double support while |sin| < 0.1."""
    s = torch.sin(2 * torch.pi * gait_phase(C, S))
    m = torch.zeros((S.N, 2), device=C.device)
    m[:, 0] = s >= 0
    m[:, 1] = s < 0
    m[torch.abs(s) < 0.1] = 1
    return m


def reference_pose(C, S, sim):
    """t1:250-274: sinusoidal hip-pitch / knee / ankle reference for the swing leg."""
    q = sim_views(sim, S.N, C.num_dof, C.num_bodies)[0]
    s = torch.sin(2 * torch.pi * gait_phase(C, S))
    sl, sr = s.clone(), s.clone()
    a1 = C.cfg.rewards.target_joint_pos_scale
    a2 = 2 * a1
    ref = torch.zeros_like(q)
    sl[sl > 0] = 0
    ref[:, 2], ref[:, 3], ref[:, 4] = sl * a1, -sl * a2, sl * a1
    sr[sr < 0] = 0
    ref[:, 8], ref[:, 9], ref[:, 10] = -sr * a1, sr * a2, -sr * a1
    ref[torch.abs(s) < 0.1] = 0
    S.ref_action = 2 * ref
    S.ref_dof_pos = ref + C.default_dof_pos


# ------------------------------------------------------------------------------------------
# command schedule (t1:109-177)
# ------------------------------------------------------------------------------------------


def resample_commands(C, S, u_cmd):
    """t1:126-177.  `u_cmd` is (3 gaits, N, 3): uniforms for (x, y, yaw) of each gait slot."""
    cr = S.command_ranges
    for g, kind in enumerate(C.cfg.commands.gait):
        ids = (S.episode_length_buf == S.gait_time[:, g]).nonzero(as_tuple=False).flatten()
        if len(ids) == 0:
            continue
        zero = torch.zeros(len(ids), device=C.device)
        draw = lambda key, col: _affine(cr[key][0], cr[key][1], u_cmd[g, ids, col:col + 1]).squeeze(1)
        assert kind in GAIT_KINDS
        S.commands[ids, 0] = draw("lin_vel_x", 0) if kind in ("walk_sagittal", "walk_omnidirectional") else zero
        S.commands[ids, 1] = draw("lin_vel_y", 1) if kind in ("walk_lateral", "walk_omnidirectional") else zero
        # t1:141-176: in heading mode the third draw is the heading target (column 3); the yaw rate (column 2) is then
        # not touched here but recomputed from the heading error for every env, every step (t1:185-188)
        if C.heading_command:
            S.commands[ids, 3] = draw("heading", 2) if kind in ("rotate", "walk_omnidirectional") else zero
        else:
            S.commands[ids, 2] = draw("ang_vel_yaw", 2) if kind in ("rotate", "walk_omnidirectional") else zero


def wrap_to_pi(angles):
    """humanoid/utils/math.py:15-18 (in place on its argument, like the reference)."""
    angles %= 2 * np.pi
    angles -= 2 * np.pi * (angles > np.pi)
    return angles


def heading_to_yaw_rate(C, S):
    """t1:185-188."""
    forward = quat_apply(S.base_quat, C.forward_vec.expand(S.N, 3))
    heading = torch.atan2(forward[:, 1], forward[:, 0])
    S.commands[:, 2] = torch.clip(0.5 * wrap_to_pi(S.commands[:, 3] - heading), -1., 1.)


def generate_gait_time(C, S, ids, u_gait):
    """t1:109-124: split the episode into the gait schedule; int32 truncation of the
    cumulative start steps.  (`np.float64 / Tensor` dispatches to `Tensor.__rtruediv__`,
    i.e. reciprocal-then-multiply — kept by writing the same expression.)"""
    if len(ids) == 0:
        return
    cols = []
    for g, kind in enumerate(C.cfg.commands.gait):
        lo, hi = C.cfg.commands.gait_time_range[kind]
        cols.append(_affine(lo, hi, u_gait[ids, g:g + 1]))
    r = torch.cat(cols, dim=1)
    scaled = r * (C.max_episode_length / torch.sum(r, dim=1, keepdim=True))
    scaled[:, 1:] = scaled[:, :-1].clone()
    scaled[:, 0] *= 0.0
    S.gait_time[ids] = torch.cumsum(scaled, dim=1).int()


# ------------------------------------------------------------------------------------------
# disturbance windows (t1:193-247)
# ------------------------------------------------------------------------------------------


def _window(C, counter, update_step, durations, interval):
    i = min(int(counter / update_step), len(durations) - 1)
    return counter % interval <= durations[i] / C.dt


def push_robots(C, S, sim, u_push):
    """t1:193-203, 217-231: inside the push window every step draws a new base velocity
    (`is_first_push` is never cleared in T1, A30)."""
    dr = C.cfg.domain_rand
    if _window(C, S.common_step_counter, dr.update_step, dr.push_duration, C.push_interval):
        if S.is_first_push:
            S.rand_push_force[:, :2] = _affine(-dr.max_push_vel_xy, dr.max_push_vel_xy, u_push[:, 0:2])
            S.rand_push_torque = _affine(-dr.max_push_ang_vel, dr.max_push_ang_vel, u_push[:, 2:5])
        sim.root_states[:, 7:9] = S.rand_push_force[:, :2]
        sim.root_states[:, 10:13] = S.rand_push_torque
    else:
        S.rand_push_force.zero_()
        S.rand_push_torque.zero_()
        S.is_first_push = True


def ext_force(C, S, u_ext):
    """t1:205-215, 233-247: draw once at window start; afterwards apply to the base of
    standing envs.  Returns the (N,3) force / torque handed to the simulator for body 0."""
    dr = C.cfg.domain_rand
    f_apply = torch.zeros(S.N, 3, device=C.device)
    t_apply = torch.zeros(S.N, 3, device=C.device)
    if _window(C, S.common_step_counter, dr.add_update_step, dr.add_duration, C.ext_force_interval):
        if S.is_first_add_force:
            fx = _affine(-dr.ext_force_max_x / 2, dr.ext_force_max_x, u_ext[:, 0:1])
            fy = _affine(-dr.ext_force_max_y, dr.ext_force_max_y, u_ext[:, 1:2])
            fz = _affine(-dr.ext_force_max_z, dr.ext_force_max_z, u_ext[:, 2:3])
            S.ext_forces = torch.cat((fx, fy, fz), 1)
            S.ext_torques = _affine(-dr.ext_torque_max, dr.ext_torque_max, u_ext[:, 3:6])
        else:
            stand = stand_command(C, S).unsqueeze(-1)
            f_apply = S.ext_forces * stand
            t_apply = S.ext_torques * stand
        S.is_first_add_force = False
    else:
        S.ext_forces.zero_()
        S.ext_torques.zero_()
        S.is_first_add_force = True
    S.applied_force, S.applied_torque = f_apply, t_apply


# ------------------------------------------------------------------------------------------
# terrain heights (lr:1551-1587)
# ------------------------------------------------------------------------------------------


def height_coords(C, S, sim):
    """lr:1569-1575 up to (not including) the `.long()` truncation: the (N, npts, 3) fp32 grid coordinates of the scan
    points.  Separate so that a test can tell a sample that legitimately sits on a cell edge from a wrong cell."""
    t = C.cfg.terrain
    npts = C.height_points.shape[0]
    pts = quat_apply_yaw(S.base_quat.repeat(1, npts), C.height_points.unsqueeze(0).expand(S.N, -1, -1).contiguous())
    pts = pts + sim.root_states[:, :3].unsqueeze(1)
    pts += t.border_size
    return pts / t.horizontal_scale


def sample_heights(C, S, sim, height_samples):
    t = C.cfg.terrain
    if t.mesh_type == "plane":
        return torch.zeros(S.N, C.height_points.shape[0], device=C.device)
    pts = height_coords(C, S, sim).long()
    px = torch.clip(pts[:, :, 0].reshape(-1), 0, height_samples.shape[0] - 2)
    py = torch.clip(pts[:, :, 1].reshape(-1), 0, height_samples.shape[1] - 2)
    h = torch.min(torch.min(height_samples[px, py], height_samples[px + 1, py]), height_samples[px, py + 1])
    return h.view(S.N, -1) * t.vertical_scale


# ------------------------------------------------------------------------------------------
# rewards (t1:572-946), evaluated in alphabetical order (A1)
# ------------------------------------------------------------------------------------------


def _pair_distance_reward(xy, lo, hi):
    d = torch.norm(xy[:, 0, :] - xy[:, 1, :], dim=1)
    near = torch.clamp(d - lo, -0.5, 0)
    far = torch.clamp(d - hi, 0, 0.5)
    return (torch.exp(-torch.abs(near) * 100) + torch.exp(-torch.abs(far) * 100)) / 2


def reward_terms(C, S, sim):
    """name -> callable returning the unscaled (N,) term.  Several terms mutate state
    (feet_air_time, feet_clearance, the phase side effect), exactly where the reference does."""
    rw, cm = C.cfg.rewards, C.cfg.commands
    q, qd, cf, rs = sim_views(sim, S.N, C.num_dof, C.num_bodies)
    root = sim.root_states
    feet_contact = lambda: cf[:, C.feet, 2] > 5.

    def action_smoothness():                                               # t1:877-892
        w = torch.ones(1, C.num_dof, device=C.device)
        d1 = (S.last_actions - S.actions) * w
        d2 = (S.actions + S.last_last_actions - 2 * S.last_actions) * w
        d3 = S.actions * w
        return (torch.sum(torch.square(d1), dim=1) + torch.sum(torch.square(d2), dim=1)
                + 0.05 * torch.sum(torch.abs(d3), dim=1))

    def base_acc():                                                        # t1:717-724
        return torch.exp(-torch.norm(S.last_root_vel - root[:, 7:13], dim=1) * 3)

    def base_height():                                                     # t1:706-715
        m = stance_mask(C, S)
        ground = torch.sum(rs[:, C.feet, 2] * m, dim=1) / torch.sum(m, dim=1)
        h = root[:, 2] - (ground - 0.05)
        return torch.exp(-torch.abs(h - rw.base_height_target) * 100)

    def collision():                                                       # t1:870-875
        return torch.sum(1. * (torch.norm(cf[:, C.pen_bodies, :], dim=-1) > 0.1), dim=1)

    def default_joint_pos():                                               # t1:686-703
        d = q - C.default_dof_pos
        yr = torch.norm(d[:, [0, 1, 5]], dim=1) + torch.norm(d[:, [6, 7, 11]], dim=1)
        yr = torch.clamp(yr - 0.1, 0, 50)
        return torch.exp(-yr * 100) - 0.01 * torch.norm(d, dim=1)

    def dof_acc():                                                         # t1:863-868
        return torch.sum(torch.square((S.last_dof_vel - qd) / C.dt), dim=1)

    def dof_vel():                                                         # t1:856-861
        return torch.sum(torch.square(qd), dim=1)

    def feet_air_time():                                                   # t1:642-657 (A6, A8)
        contact = feet_contact()
        m = stance_mask(C, S).clone()
        m[torch.norm(S.commands[:, :3], dim=1) < 0.05] = 1
        S.contact_filt = torch.logical_or(torch.logical_or(contact, m), S.last_contacts)
        S.last_contacts = contact
        first = (S.feet_air_time > 0.) * S.contact_filt
        S.feet_air_time += C.dt
        r = S.feet_air_time.clamp(0, 0.5) * first
        S.feet_air_time *= ~S.contact_filt
        return r.sum(dim=1)

    def feet_clearance():                                                  # t1:793-814 (A9)
        contact = feet_contact()
        z = rs[:, C.feet, 2]
        S.feet_height += z - S.last_feet_z
        S.last_feet_z = z.clone()      # the reference keeps a view; with refreshed sim tensors a copy is equivalent
        swing = 1 - stance_mask(C, S)
        hit = (S.feet_height > rw.target_feet_height) * (S.feet_height < rw.target_feet_height_max)
        r = torch.sum(hit * swing, dim=1)
        S.feet_height *= ~contact
        return r

    def feet_contact_forces():                                             # t1:679-684
        return torch.sum((torch.norm(cf[:, C.feet, :], dim=-1) - rw.max_contact_force).clip(0, 400), dim=1)

    def feet_contact_number():                                             # t1:659-668 (A14)
        contact = feet_contact()
        m = stance_mask(C, S).clone()
        m[torch.norm(S.commands[:, :3], dim=1) <= cm.stand_com_threshold] = 1
        return torch.mean(torch.where(contact == m, 1, -0.3), dim=1)

    def feet_distance():                                                   # t1:599-612
        return _pair_distance_reward(rs[:, C.feet, :2], rw.foot_min_dist, rw.foot_max_dist)

    def feet_rotation():                                                   # t1:926-935 (A11)
        rot = torch.sum(torch.square(S.feet_euler_xyz[:, :, 1]), dim=1)
        return 1 * torch.exp(-(rot / 1) ** 2)

    def foot_slip():                                                       # t1:630-640 (A10)
        r = torch.sqrt(torch.norm(rs[:, C.feet, 10:12], dim=2))
        r *= feet_contact()
        return torch.sum(r, dim=1)

    def joint_pos():                                                       # t1:576-596 (A3)
        target = S.ref_dof_pos.clone()
        stand = stand_command(C, S)
        target[stand] = C.default_dof_pos.clone()
        n = torch.norm(q.clone() - target, dim=1)
        r = torch.exp(-2 * n) - 0.2 * n.clamp(0, 0.5)
        r[stand] = 1.0
        return r

    def knee_distance():                                                   # t1:615-628
        return _pair_distance_reward(rs[:, C.knees, :2], rw.knee_min_dist, rw.knee_max_dist)

    def low_speed():                                                       # t1:816-847 (A13)
        v, c = S.base_lin_vel[:, 0], S.commands[:, 0]
        slow = torch.abs(v) < 0.5 * torch.abs(c)
        fast = torch.abs(v) > 1.2 * torch.abs(c)
        r = torch.zeros_like(v)
        r[slow] = -1.0
        r[fast] = 0
        r[~(slow | fast)] = 1.2
        r[torch.sign(v) != torch.sign(c)] = -2.0
        return r * (c.abs() > 0.05)

    def orientation():                                                     # t1:670-677
        a = torch.exp(-torch.sum(torch.abs(S.base_euler_xyz[:, :2]), dim=1) * 10)
        b = torch.exp(-torch.norm(S.projected_gravity[:, :2], dim=1) * 20)
        return (a + b) / 2.

    def stand_still():                                                     # t1:899-915 (A12)
        idx = [0, 1, 2, 3, 5, 6, 7, 8]
        w = torch.tensor([[2.0, 2.0, 1.0, 1.0, 1.0, 2.0, 2.0, 1.0, 1.0, 1.0]], device=C.device)
        err = torch.cat((q[:, idx] - C.default_dof_pos[:, idx], S.feet_euler_xyz[:, :, 1]), dim=1) * w
        r = torch.exp(-torch.sum(torch.square(err), dim=1))
        return torch.where(stand_command(C, S), r, torch.zeros_like(r))

    def torques():                                                         # t1:849-854
        return torch.sum(torch.square(S.torques), dim=1)

    def track_vel_hard():                                                  # t1:738-758
        le = torch.norm(S.commands[:, :2] - S.base_lin_vel[:, :2], dim=1)
        ae = torch.abs(S.commands[:, 2] - S.base_ang_vel[:, 2])
        return (torch.exp(-le * 10) + torch.exp(-ae * 10)) / 2. - 0.2 * (le + ae)

    def feet_stumble():                                                    # t1:937-940 (inactive in t1_cfg)
        return torch.any(torch.norm(cf[:, C.feet, :2], dim=2) > 5 * torch.abs(cf[:, C.feet, 2]), dim=1)

    def stand_sysmetry():                                                  # t1:917-925 (inactive in t1_cfg)
        err = q[:, [0, 1, 2, 3]] - q[:, [5, 6, 7, 8]]
        r = torch.exp(-torch.sum(torch.square(err), dim=1))
        return torch.where(stand_command(C, S), r, torch.zeros_like(r))

    def tracking_ang_vel():                                                # t1:776-790
        e = S.commands[:, 2] - S.base_ang_vel[:, 2]
        return torch.where(stand_command(C, S), torch.exp(-torch.abs(e) * rw.tracking_sigma * 2),
                           torch.exp(-torch.square(e) * rw.tracking_sigma))

    def tracking_lin_vel():                                                # t1:760-774
        e = S.commands[:, :2] - S.base_lin_vel[:, :2]
        return torch.where(stand_command(C, S),
                           torch.exp(-torch.sum(torch.abs(e), dim=1) * rw.tracking_sigma * 2),
                           torch.exp(-torch.sum(torch.square(e), dim=1) * rw.tracking_sigma))

    def vel_mismatch_exp():                                                # t1:726-736
        a = torch.exp(-torch.square(S.base_lin_vel[:, 2]) * 10)
        b = torch.exp(-torch.norm(S.base_ang_vel[:, :2], dim=1) * 5.)
        return (a + b) / 2.

    return {f.__name__: f for f in (
        action_smoothness, base_acc, base_height, collision, default_joint_pos, dof_acc, dof_vel, feet_air_time,
        feet_clearance, feet_contact_forces, feet_contact_number, feet_distance, feet_rotation, feet_stumble, foot_slip,
        joint_pos, knee_distance, low_speed, orientation, stand_still, stand_sysmetry, torques, track_vel_hard,
        tracking_ang_vel, tracking_lin_vel, vel_mismatch_exp)}


def compute_reward(C, S, sim):
    """lr:654-680: sum of scaled terms in `reward_names` order, per-term episode sums,
    clip at zero.  Returns the per-term scaled values (test handle)."""
    terms = reward_terms(C, S, sim)
    S.rew_buf[:] = 0
    out = {}
    for name in C.reward_names:
        r = terms[name]() * C.reward_scales[name]
        S.rew_buf += r
        S.episode_sums[name] += r
        out[name] = r
    if C.cfg.rewards.only_positive_rewards:
        S.rew_buf[:] = torch.clip(S.rew_buf[:], min=0)
    if "termination" in C.reward_scales:                                   # lr:677-680, t1:894-896: after the clip
        r = (S.reset_buf * ~S.time_out_buf) * C.reward_scales["termination"]
        S.rew_buf += r
        S.episode_sums["termination"] += r
        out["termination"] = r
    return out


# ------------------------------------------------------------------------------------------
# reset (t1:483-559 with lr:604-651, 732-783, 1076-1169)
# ------------------------------------------------------------------------------------------


def reset_envs(C, S, sim, ids, R, terrain=None):
    if len(ids) == 0:
        return
    cfg, dr = C.cfg, C.cfg.domain_rand
    N, D = S.N, C.num_dof
    dof = sim.dof_state.view(N, D, 2)
    if C.curriculum:                                                       # lr:1138-1158
        if S.init_done:
            dist = torch.norm(sim.root_states[ids, :2] - S.env_origins[ids, :2], dim=1)
            up = dist > terrain.env_length / 2
            down = (dist < torch.norm(S.commands[ids, :2], dim=1) * C.max_episode_length_s * 0.5) * ~up
            S.terrain_levels[ids] += 1 * up - 1 * down
            S.terrain_levels[ids] = torch.where(S.terrain_levels[ids] >= terrain.max_level,
                                                R["terrain_level"][ids] % terrain.max_level,
                                                torch.clip(S.terrain_levels[ids], 0))
            S.env_origins[ids] = terrain.origins[S.terrain_levels[ids], S.terrain_types[ids]]
    if cfg.commands.curriculum and (S.common_step_counter % C.max_episode_length == 0):   # lr:1160-1169
        if torch.mean(S.episode_sums["tracking_lin_vel"][ids]) / C.max_episode_length > 0.8 * C.reward_scales["tracking_lin_vel"]:
            rx = S.command_ranges["lin_vel_x"]
            rx[0] = np.clip(rx[0] - 0.25, -cfg.commands.max_curriculum / 2, 0.)
            rx[1] = np.clip(rx[1] + 0.5, 0., cfg.commands.max_curriculum)
    # lr:1076-1090 joint state
    dof[ids, :, 0] = C.default_dof_pos + _affine(-0.1, 0.1, R["dofs"][ids])
    dof[ids, :, 1] = 0.
    # lr:1092-1120 root state
    sim.root_states[ids] = C.base_init_state
    sim.root_states[ids, :3] += S.env_origins[ids]
    if C.custom_origins:
        half = cfg.terrain.platform / 3 if cfg.terrain.curriculum else cfg.terrain.terrain_length / 2
        sim.root_states[ids, :2] += _affine(-half, half, R["root_xy"][ids])
    # lr:732-783 actuator randomisation
    u = R["dr"][ids]
    if dr.randomize_torque:
        S.torque_multi[ids] = _affine(*dr.torque_multiplier_range, u[:, 0])
    if dr.randomize_motor_offset:
        S.motor_offsets[ids, :] = _affine(*dr.motor_offset_range, u[:, 1])
    if dr.randomize_gains:
        S.randomized_p_gains[ids] = _affine(*dr.stiffness_multiplier_range, u[:, 2]) * C.p_gains
        S.randomized_d_gains[ids] = _affine(*dr.damping_multiplier_range, u[:, 3]) * C.d_gains
    if dr.randomize_coulomb_friction:
        S.randomized_joint_coulomb[ids] = _affine(*dr.joint_coulomb_range, u[:, 4])
        S.randomized_joint_viscous[ids] = _affine(*dr.joint_viscous_range, u[:, 5])
    if C.joint_props[0]:                                                   # lr:762-763: one multiplier per env
        S.joint_friction_coeffs[ids] = _affine(*dr.joint_friction_range, R["dr_joint"][ids, 0:1])
    if C.joint_props[1]:                                                   # lr:772-773
        S.joint_damping_coeffs[ids] = _affine(*dr.joint_damping_range, R["dr_joint"][ids, 1:2])
    if dr.randomize_joint_armature:
        assert dr.randomize_joint_armature_each_joint
        for j in range(D):
            S.joint_armatures[ids, j] = _affine(*getattr(dr, f"joint_{j + 1}_armature_range"), u[:, 6, j:j + 1]).reshape(-1)
    # lr:604-633 lag buffers and indices
    # (without randomize_*_lag_timesteps the index is the range maximum, and the reference keeps it as a float tensor)
    li = R["lag_idx"][ids]
    if dr.add_lag:
        S.lag_buffer[ids, :, :] = 0.0
        S.lag_timestep[ids] = li[:, 0] if dr.randomize_lag_timesteps else dr.lag_timesteps_range[1]
    if dr.add_dof_lag:
        S.dof_lag_buffer[ids, :, :] = 0.0
        S.dof_lag_timestep[ids] = li[:, 1] if dr.randomize_dof_lag_timesteps else dr.dof_lag_timesteps_range[1]
    if dr.add_imu_lag:
        S.imu_lag_buffer[ids, :, :] = 0.0
        S.imu_lag_timestep[ids] = li[:, 2] if dr.randomize_imu_lag_timesteps else dr.imu_lag_timesteps_range[1]
    if C.pos_vel_lag:                                                      # lr:634-650
        pv = R["lag_idx_pv"][ids]
        S.dof_pos_lag_buffer[ids, :, :] = 0.0
        S.dof_vel_lag_buffer[ids, :, :] = 0.0
        S.dof_pos_lag_timestep[ids] = pv[:, 0] if dr.randomize_dof_pos_lag_timesteps else dr.dof_pos_lag_timesteps_range[1]
        S.dof_vel_lag_timestep[ids] = pv[:, 1] if dr.randomize_dof_vel_lag_timesteps else dr.dof_vel_lag_timesteps_range[1]
    for k, on in (("lag", dr.add_lag), ("dof_lag", dr.add_dof_lag), ("imu_lag", dr.add_imu_lag),
                  ("dof_pos_lag", C.pos_vel_lag), ("dof_vel_lag", C.pos_vel_lag)):
        if on and C.perstep[k]:                                            # lr:610-611, 620-621, 630-631, 641-642, 648-649
            getattr(S, f"last_{k}_timestep")[ids] = getattr(dr, f"{k}_timesteps_range")[1]
    # t1:513-523
    for name in ("last_last_actions", "actions", "last_actions", "last_dof_vel", "last_root_vel", "feet_air_time"):
        getattr(S, name)[ids] = 0.
    S.episode_length_buf[ids] = 0
    S.phase_length_buf[ids] = 0
    S.reset_buf[ids] = 1
    S.gait_start[ids] = R["gait_start"][ids] * 0.5
    generate_gait_time(C, S, ids, R["gait_time"])
    resample_commands(C, S, R["cmd"][1])                                   # all envs again (A24)
    # t1:530-541 episode statistics
    S.extras["episode"] = {}
    for key in S.episode_sums:
        S.extras["episode"]["rew_" + key] = torch.mean(S.episode_sums[key][ids]) / C.max_episode_length_s
        S.episode_sums[key][ids] = 0.
    if cfg.terrain.mesh_type == "trimesh":
        S.extras["episode"]["terrain_level"] = torch.mean(S.terrain_levels.float())
    if cfg.commands.curriculum:
        S.extras["episode"]["max_command_x"] = S.command_ranges["lin_vel_x"][1]
    if cfg.env.send_timeouts:
        S.extras["time_outs"] = S.time_out_buf
    # t1:548-554 derived base state of the re-spawned robots
    _, _, _, rs = sim_views(sim, N, D, C.num_bodies)
    S.base_quat[ids] = sim.root_states[ids, 3:7]
    S.base_euler_xyz = euler_xyz(S.base_quat)
    S.projected_gravity[ids] = quat_rotate_inverse(S.base_quat[ids], C.gravity_vec.expand(len(ids), 3))
    S.base_lin_vel[ids] = quat_rotate_inverse(S.base_quat[ids], sim.root_states[ids, 7:10])
    S.base_ang_vel[ids] = quat_rotate_inverse(S.base_quat[ids], sim.root_states[ids, 10:13])
    S.feet_euler_xyz = euler_xyz(rs[:, C.feet, 3:7])
    # t1:556-559: `*= 0` keeps the sign bit / NaNs of the old entries (A19)
    S.obs_history[:, ids] *= 0
    S.critic_history[:, ids] *= 0


# ------------------------------------------------------------------------------------------
# observations (t1:368-481)
# ------------------------------------------------------------------------------------------


def compute_observations(C, S, sim, u_noise, lag_step=None):
    cfg, dr, os_ = C.cfg, C.cfg.domain_rand, C.obs_scales
    N, D = S.N, C.num_dof
    q, qd, cf, _ = sim_views(sim, N, D, C.num_bodies)
    ar = torch.arange(N, device=C.device)
    phase = gait_phase(C, S)
    reference_pose(C, S, sim)
    sin_p = torch.sin(2 * torch.pi * phase).unsqueeze(1)
    cos_p = torch.cos(2 * torch.pi * phase).unsqueeze(1)
    stance = stance_mask(C, S)
    contact = cf[:, C.feet, 2] > 5
    S.command_input = torch.cat((sin_p, cos_p, S.commands[:, :3] * C.commands_scale), dim=1)
    push_f, push_t = S.rand_push_force[:, :2], S.rand_push_torque
    if dr.add_ext_force:
        push_f = S.ext_forces[:, :2] / (dr.ext_force_max_x + 0.1)
        push_t = S.ext_torques / (dr.ext_torque_max + 0.1)
    priv = torch.cat((
        S.command_input, (q - C.default_dof_pos) * os_.dof_pos, qd * os_.dof_vel, S.actions, q - S.ref_dof_pos,
        S.base_lin_vel * os_.lin_vel, S.base_ang_vel * os_.ang_vel, S.base_euler_xyz * os_.quat,
        push_f[:, :2], push_t, S.env_frictions, S.body_mass / 30., stance, contact), dim=-1)
    DEC = C.decimation                                                     # rows of `lag_step` behind the substeps'
    if dr.add_dof_lag:
        if C.perstep["dof_lag"]:                                           # t1:408-413
            _redraw_lag(S, "dof_lag", lag_step[DEC])
        S.lagged_dof_pos = S.dof_lag_buffer[ar, :D, S.dof_lag_timestep.int()]
        S.lagged_dof_vel = S.dof_lag_buffer[ar, -D:, S.dof_lag_timestep.int()]
    elif C.pos_vel_lag:                                                    # t1:416-431
        if C.perstep["dof_pos_lag"]:
            _redraw_lag(S, "dof_pos_lag", lag_step[DEC + 2])
        S.lagged_dof_pos = S.dof_pos_lag_buffer[ar, :, S.dof_pos_lag_timestep.int()]
        if C.perstep["dof_vel_lag"]:
            _redraw_lag(S, "dof_vel_lag", lag_step[DEC + 3])
        S.lagged_dof_vel = S.dof_vel_lag_buffer[ar, :, S.dof_vel_lag_timestep.int()]
    else:
        S.lagged_dof_pos, S.lagged_dof_vel = q, qd
    if dr.add_imu_lag:
        if C.perstep["imu_lag"]:                                           # t1:437-442
            _redraw_lag(S, "imu_lag", lag_step[DEC + 1])
        imu = S.imu_lag_buffer[ar, :, S.imu_lag_timestep.int()]
        S.lagged_base_ang_vel, S.lagged_base_euler_xyz = imu[:, :3].clone(), imu[:, -3:].clone()
    else:
        S.lagged_base_ang_vel, S.lagged_base_euler_xyz = S.base_ang_vel[:, :3], S.base_euler_xyz[:, -3:]
    frame = torch.cat((
        S.command_input, (S.lagged_dof_pos - C.default_dof_pos) * os_.dof_pos, S.lagged_dof_vel * os_.dof_vel,
        S.actions, S.lagged_base_ang_vel * os_.ang_vel, S.lagged_base_euler_xyz * os_.quat), dim=-1)
    if cfg.terrain.measure_heights:
        h = torch.clip(sim.root_states[:, 2].unsqueeze(1) - 0.5 - S.measured_heights, -1, 1.) * os_.height_measurements
        priv = torch.cat((priv.clone(), h), dim=-1)
    if cfg.noise.add_noise:
        frame = frame.clone() + (2 * u_noise - 1) * C.noise_scale_vec * cfg.noise.noise_level
    S.obs_history = torch.cat((S.obs_history[1:], frame.unsqueeze(0)), dim=0)
    S.critic_history = torch.cat((S.critic_history[1:], priv.unsqueeze(0)), dim=0)
    S.obs_buf = S.obs_history.permute(1, 0, 2).reshape(N, -1)
    S.privileged_obs_buf = S.critic_history.permute(1, 0, 2).reshape(N, -1)


# ------------------------------------------------------------------------------------------
# the step (lr:387-506, t1:179-215, 360-366)
# ------------------------------------------------------------------------------------------


def post_physics(C, S, sim, R, terrain=None, height_samples=None):
    cfg, dr = C.cfg, C.cfg.domain_rand
    N, D = S.N, C.num_dof
    q, qd, cf, rs = sim_views(sim, N, D, C.num_bodies)
    S.episode_length_buf += 1
    S.common_step_counter += 1
    S.base_quat[:] = sim.root_states[:, 3:7]                               # lr:475-481
    S.base_lin_vel[:] = quat_rotate_inverse(S.base_quat, sim.root_states[:, 7:10])
    S.base_ang_vel[:] = quat_rotate_inverse(S.base_quat, sim.root_states[:, 10:13])
    S.projected_gravity[:] = quat_rotate_inverse(S.base_quat, C.gravity_vec.expand(N, 3))
    S.base_euler_xyz = euler_xyz(S.base_quat)
    S.feet_euler_xyz = euler_xyz(rs[:, C.feet, 3:7])
    S.phase_length_buf += 1                                                # t1:183-215
    resample_commands(C, S, R["cmd"][0])
    if C.heading_command:
        heading_to_yaw_rate(C, S)
    if cfg.terrain.measure_heights:
        S.measured_heights = sample_heights(C, S, sim, height_samples)
    if dr.push_robots:
        push_robots(C, S, sim, R["push"])
    if dr.add_ext_force:
        ext_force(C, S, R["ext"])
    # lr:509-517
    S.reset_buf = torch.any(torch.norm(cf[:, C.term_bodies, :], dim=-1) > 1, dim=1)
    S.time_out_buf = S.episode_length_buf > C.max_episode_length
    S.reset_buf |= S.time_out_buf
    S.reward_terms = compute_reward(C, S, sim)
    S.reset_ids = S.reset_buf.nonzero(as_tuple=False).flatten()
    reset_envs(C, S, sim, S.reset_ids, R, terrain)
    compute_observations(C, S, sim, R["noise"], R.get("lag_step"))
    S.last_last_actions[:] = S.last_actions                                # lr:496-499 (live state only)
    S.last_actions[:] = S.actions
    S.last_dof_vel[:] = qd
    S.last_root_vel[:] = sim.root_states[:, 7:13]


def step(C, S, sim, actions, R, terrain=None, height_samples=None, physics=None):
    """One policy step.  `physics(substep)` stands in for gym.simulate + refresh."""
    clip_a = C.cfg.normalization.clip_actions
    if getattr(C.cfg.env, "use_ref_actions", False):                        # t1:360-366 (in place on the caller's tensor)
        actions += S.ref_action
    S.actions = torch.clip(actions, -clip_a, clip_a).to(C.device)
    for k in range(C.decimation):
        S.torques = torque_substep(C, S, sim, S.actions, R["torque"][k],
                                   R["lag_step"][k] if C.perstep["lag"] else None).view(S.torques.shape)
        if physics is not None:
            physics(k)
        lag_push(C, S, sim)
    post_physics(C, S, sim, R, terrain, height_samples)
    clip_o = C.cfg.normalization.clip_observations
    S.obs_buf = torch.clip(S.obs_buf, -clip_o, clip_o)
    S.privileged_obs_buf = torch.clip(S.privileged_obs_buf, -clip_o, clip_o)
    return S.obs_buf, S.privileged_obs_buf, S.rew_buf, S.reset_buf, S.extras


# ------------------------------------------------------------------------------------------
# GAE (rollout_storage.py:97-119)
# ------------------------------------------------------------------------------------------


def gae_returns(rewards, values, dones, last_values, gamma, lam):
    """Reverse scan over T for (T,N,1) fp32 rewards/values and uint8 dones; returns
    (returns, normalised advantages) with the unbiased std (A27)."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    adv = 0
    for t in reversed(range(T)):
        nxt = last_values if t == T - 1 else values[t + 1]
        alive = 1.0 - dones[t].float()
        delta = rewards[t] + alive * gamma * nxt - values[t]
        adv = delta + alive * gamma * lam * adv
        returns[t] = adv + values[t]
    a = returns - values
    return returns, (a - a.mean()) / (a.std() + 1e-8)


# ------------------------------------------------------------------------------------------
# state exchange with the fixtures / the CUDA env (flat dict, reference attribute names)
# ------------------------------------------------------------------------------------------

_PLAIN_STATE = ("torques actions last_actions last_last_actions last_dof_vel last_root_vel commands feet_air_time "
                "feet_height last_contacts contact_filt base_quat base_lin_vel base_ang_vel projected_gravity "
                "base_euler_xyz feet_euler_xyz ext_forces ext_torques rand_push_force rand_push_torque ref_dof_pos "
                "gait_time gait_start torque_multi motor_offsets randomized_p_gains randomized_d_gains "
                "randomized_joint_coulomb randomized_joint_viscous joint_armatures lag_buffer dof_lag_buffer "
                "imu_lag_buffer lag_timestep dof_lag_timestep imu_lag_timestep episode_length_buf phase_length_buf "
                "rew_buf reset_buf time_out_buf env_origins env_frictions body_mass").split()


OPTIONAL_LAG_STATE = ("last_lag_timestep last_dof_lag_timestep last_imu_lag_timestep dof_pos_lag_buffer dof_vel_lag_buffer dof_pos_lag_timestep dof_vel_lag_timestep last_dof_pos_lag_timestep last_dof_vel_lag_timestep joint_friction_coeffs joint_damping_coeffs").split()


def load_state(C, S, state):
    """Adopt a flat state dict (the `state0.*` entries of tests/golden/*.npz)."""
    for k in _PLAIN_STATE:
        setattr(S, k, torch.as_tensor(state[k]).clone().to(C.device))
    if "ref_action" in state:          # fixtures written since env.use_ref_actions is pinned carry it
        S.ref_action = torch.as_tensor(state["ref_action"]).clone().to(C.device)
    for k in OPTIONAL_LAG_STATE:       # ... and the state of the per-step / position-velocity lag options
        if k in state:
            setattr(S, k, torch.as_tensor(state[k]).clone().to(C.device))
    S.last_feet_z = torch.as_tensor(state["last_feet_z"]).clone().to(C.device)
    S.obs_history = torch.as_tensor(state["obs_history"]).clone().to(C.device)
    S.critic_history = torch.as_tensor(state["critic_history"]).clone().to(C.device)
    es = torch.as_tensor(state["episode_sums"]).to(C.device)
    S.episode_sums = {k: es[i].clone() for i, k in enumerate(C.reward_scales)}
    cnt = [int(v) for v in state["counters"]]
    S.common_step_counter, S.is_first_add_force, S.is_first_push = cnt[0], bool(cnt[1]), bool(cnt[2])
    cr = torch.as_tensor(state["command_ranges"]).tolist()
    S.command_ranges = dict(S.command_ranges)
    for i, k in enumerate(("lin_vel_x", "lin_vel_y", "ang_vel_yaw")):
        S.command_ranges[k] = list(cr[i])
    if "terrain_levels" in state:
        S.terrain_levels = torch.as_tensor(state["terrain_levels"]).clone().to(C.device)
        S.terrain_types = torch.as_tensor(state["terrain_types"]).clone().to(C.device)
    return S
