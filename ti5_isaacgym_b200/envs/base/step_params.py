"""cfg -> `Ti5Params`: every scalar the kernels need, computed the way the reference's Python
computes it (doubles first, rounded to fp32 exactly where torch rounds).  Pure host code.

Follows `_parse_cfg` (lr:94-113), `_init_buffers` (lr:212-249), `_prepare_reward_function`
(lr:352-384), `_process_dof_props` (lr:837-849) and `_get_noise_scale_vec` (t1:326-357).
"""
import os

import numpy as np

from ... import _lib
from ...utils.helpers import class_to_dict

C = _lib.CONSTS
TERM_NAMES = ("action_smoothness", "base_acc", "base_height", "collision", "default_joint_pos", "dof_acc", "dof_vel",
              "dof_vel_limits", "feet_air_time", "feet_clearance", "feet_contact_forces", "feet_contact_number",
              "feet_distance", "feet_rotation", "feet_stumble", "foot_slip", "joint_pos", "knee_distance", "low_speed",
              "orientation", "stand_still", "stand_sysmetry", "termination", "torques", "track_vel_hard",
              "tracking_ang_vel", "tracking_lin_vel", "vel_mismatch_exp")
GAIT_KIND = {"stand": C["TI5_GAIT_STAND"], "walk_sagittal": C["TI5_GAIT_WALK_SAGITTAL"],
             "walk_lateral": C["TI5_GAIT_WALK_LATERAL"], "rotate": C["TI5_GAIT_ROTATE"],
             "walk_omnidirectional": C["TI5_GAIT_WALK_OMNI"]}
# lr:755-773 `_each_joint`: t1_cfg has per-joint ranges for ten of the twelve joints only (the reference itself raises
# AttributeError at joint 11); the one-multiplier-per-env forms are supported
UNSUPPORTED_FLAGS = ("randomize_joint_friction_each_joint", "randomize_joint_damping_each_joint")


def pick_env_block(num_envs, sms=148):
    """Envs per CTA of the per-env kernels.  32 while the grid is at most two CTAs per SM (<= 9472 envs: the BASELINE
    size, 8192 envs = 256 CTAs, is in this regime — four warps per 32 envs in ti5_reset_observe, common carve-out and
    early mode in ti5_post_physics); 64 beyond.  Measured us/step, 32 vs 64: 4096 envs 47.9 / 50.0, 8192 envs 50.2 / 54.0,
    10240 envs 58.6 / 56.4, 12288 envs 64.7 / 60.2, 16384 envs 79.3 / 66.4; 128 was no better than 64 anywhere."""
    return 32 if num_envs <= 2 * sms * 32 else 64


def reward_scales(cfg, dt):
    """lr:357-364: drop zero scales, multiply the rest by dt (in double); alphabetical order."""
    return {k: v * dt for k, v in class_to_dict(cfg.rewards.scales).items() if v != 0}


def build_params(cfg, sim_dt, robot, terrain=None, height_shape=(0, 0), div_mode=None, rng_mode=None, seed=0,
                 env_block=None):
    dr, rw, cm, nz = cfg.domain_rand, cfg.rewards, cfg.commands, cfg.normalization
    for f in UNSUPPORTED_FLAGS:
        if getattr(dr, f, False):
            raise NotImplementedError(f"domain_rand.{f}=True is not exercised by t1_dh_stand (t1_cfg:290-312)")
    p = _lib.Ti5Params()
    dt = cfg.control.decimation * sim_dt
    N = cfg.env.num_envs
    measure = bool(cfg.terrain.measure_heights)
    p.num_envs, p.frame_stack, p.c_frame_stack = N, cfg.env.frame_stack, cfg.env.c_frame_stack
    p.num_single_obs = cfg.env.num_single_obs
    p.priv_frame = cfg.env.single_num_privileged_obs + (cfg.terrain.num_height if measure else 0)
    p.decimation = cfg.control.decimation
    p.lag_len = dr.lag_timesteps_range[1] + 1
    # separate position / velocity lags (t1:416-431) only show when the common joint-state lag is off; both read the DOF
    # ring, which then has to cover both ranges
    pos_vel = bool(getattr(dr, "add_dof_pos_vel_lag", False)) and not dr.add_dof_lag
    p.dof_lag_len = (max(dr.dof_pos_lag_timesteps_range[1], dr.dof_vel_lag_timesteps_range[1]) if pos_vel
                     else dr.dof_lag_timesteps_range[1]) + 1
    p.imu_lag_len = dr.imu_lag_timesteps_range[1] + 1
    custom = cfg.terrain.mesh_type in ("heightfield", "trimesh")
    curriculum = bool(cfg.terrain.curriculum) and custom                     # lr:104-105
    flag_src = {
        "TI5_F_ADD_LAG": dr.add_lag, "TI5_F_ADD_DOF_LAG": dr.add_dof_lag, "TI5_F_ADD_IMU_LAG": dr.add_imu_lag,
        "TI5_F_RAND_GAINS": dr.randomize_gains, "TI5_F_RAND_COULOMB": dr.randomize_coulomb_friction,
        "TI5_F_RAND_TORQUE": dr.randomize_torque, "TI5_F_RAND_MOTOR_OFFSET": dr.randomize_motor_offset,
        "TI5_F_RAND_ARMATURE": dr.randomize_joint_armature, "TI5_F_ADD_NOISE": cfg.noise.add_noise,
        "TI5_F_MEASURE_HEIGHTS": measure, "TI5_F_PUSH_ROBOTS": dr.push_robots, "TI5_F_ADD_EXT_FORCE": dr.add_ext_force,
        "TI5_F_ONLY_POSITIVE": rw.only_positive_rewards, "TI5_F_CUSTOM_ORIGINS": custom,
        "TI5_F_TERRAIN_CURRICULUM": curriculum, "TI5_F_COMMAND_CURRICULUM": cm.curriculum,
        "TI5_F_TRIMESH": cfg.terrain.mesh_type == "trimesh", "TI5_F_RAND_LAG_STEPS": dr.randomize_lag_timesteps,
        "TI5_F_RAND_DOF_LAG_STEPS": dr.randomize_dof_lag_timesteps,
        "TI5_F_RAND_IMU_LAG_STEPS": dr.randomize_imu_lag_timesteps, "TI5_F_PLANE": cfg.terrain.mesh_type == "plane",
        "TI5_F_HEADING_COMMAND": cm.heading_command, "TI5_F_NO_SW_SWITCH": not cm.sw_switch,
        # options t1_cfg marks "always False" (a per-step re-draw needs its lag and the randomisation of its index on)
        "TI5_F_LAG_PERSTEP": dr.add_lag and dr.randomize_lag_timesteps and getattr(dr, "randomize_lag_timesteps_perstep", False),
        "TI5_F_DOF_LAG_PERSTEP": dr.add_dof_lag and dr.randomize_dof_lag_timesteps
                                 and getattr(dr, "randomize_dof_lag_timesteps_perstep", False),
        "TI5_F_IMU_LAG_PERSTEP": dr.add_imu_lag and dr.randomize_imu_lag_timesteps
                                 and getattr(dr, "randomize_imu_lag_timesteps_perstep", False),
        "TI5_F_POS_VEL_LAG": pos_vel,
        "TI5_F_RAND_POS_LAG_STEPS": pos_vel and dr.randomize_dof_pos_lag_timesteps,
        "TI5_F_RAND_VEL_LAG_STEPS": pos_vel and dr.randomize_dof_vel_lag_timesteps,
        "TI5_F_POS_LAG_PERSTEP": pos_vel and dr.randomize_dof_pos_lag_timesteps
                                 and getattr(dr, "randomize_dof_pos_lag_timesteps_perstep", False),
        "TI5_F_VEL_LAG_PERSTEP": pos_vel and dr.randomize_dof_vel_lag_timesteps
                                 and getattr(dr, "randomize_dof_vel_lag_timesteps_perstep", False)}
    if pos_vel:
        flag_src["TI5_F_ADD_DOF_LAG"] = True          # the DOF ring is pushed (header: TI5_F_POS_VEL_LAG)
    p.flags = sum(C[k] for k, on in flag_src.items() if on)
    if dr.randomize_joint_armature and not dr.randomize_joint_armature_each_joint:
        raise NotImplementedError("randomize_joint_armature without _each_joint is not used by t1_dh_stand")
    p.div_mode = C["TI5_DIV_RECIPROCAL"] if div_mode is None else div_mode   # torch-on-GPU semantics by default
    p.rng_mode = C["TI5_RNG_PHILOX"] if rng_mode is None else rng_mode
    p.env_block = env_block or int(os.environ.get("TI5_ENV_BLOCK", 0)) or pick_env_block(N)
    p.seed = seed
    p.applied_stride = 3                    # plain (N,3) applied_force / applied_torque unless the env re-points them
    gaits = list(cm.gait)
    assert len(gaits) <= C["TI5_MAX_GAITS"]
    p.num_gaits = len(gaits)
    for i, g in enumerate(gaits):
        p.gait_kind[i] = GAIT_KIND[g]
        lo, hi = cm.gait_time_range[g]
        p.gait_time_w[i], p.gait_time_lo[i] = hi - lo, lo
    p.num_height_points = cfg.terrain.num_height if measure else 0
    p.height_rows, p.height_cols = height_shape
    for i in range(2):
        p.feet[i], p.knees[i] = robot.feet_indices[i], robot.knee_indices[i]
    assert len(robot.termination_contact_indices) == 1 and len(robot.penalised_contact_indices) == 1
    p.term_body, p.pen_body = robot.termination_contact_indices[0], robot.penalised_contact_indices[0]
    for i, rng in enumerate((dr.lag_timesteps_range, dr.dof_lag_timesteps_range, dr.imu_lag_timesteps_range)):
        p.lag_range[i][0], p.lag_range[i][1] = rng
    for i, rng in enumerate((dr.dof_pos_lag_timesteps_range, dr.dof_vel_lag_timesteps_range)):
        p.lag_range_pv[i][0], p.lag_range_pv[i][1] = rng
    p.flags2 = (C["TI5_F2_RAND_JOINT_FRICTION"] if getattr(dr, "randomize_joint_friction", False) else 0) | \
               (C["TI5_F2_RAND_JOINT_DAMPING"] if getattr(dr, "randomize_joint_damping", False) else 0)
    p.joint_friction_w, p.joint_friction_lo = dr.joint_friction_range[1] - dr.joint_friction_range[0], dr.joint_friction_range[0]
    p.joint_damping_w, p.joint_damping_lo = dr.joint_damping_range[1] - dr.joint_damping_range[0], dr.joint_damping_range[0]
    # time scales (lr:96-113)
    p.dt = dt
    p.max_episode_length_s = cfg.env.episode_length_s
    p.max_episode_length = int(np.ceil(cfg.env.episode_length_s / dt))
    p.push_interval = int(np.ceil(dr.push_interval_s / dt))
    p.ext_force_interval = int(np.ceil(dr.ext_force_interval_s / dt))
    p.push_update_step, p.add_update_step = int(dr.update_step), int(dr.add_update_step)
    assert len(dr.push_duration) <= C["TI5_MAX_WINDOWS"] and len(dr.add_duration) <= C["TI5_MAX_WINDOWS"]
    p.n_push_dur, p.n_add_dur = len(dr.push_duration), len(dr.add_duration)
    for i, d in enumerate(dr.push_duration):
        p.push_duration[i] = d / dt
    for i, d in enumerate(dr.add_duration):
        p.add_duration[i] = d / dt
    p.cycle_time, p.action_scale = rw.cycle_time, cfg.control.action_scale
    p.clip_actions, p.clip_obs = nz.clip_actions, nz.clip_observations
    p.stand_threshold = cm.stand_com_threshold
    p.heading_w, p.heading_lo = cm.ranges.heading[1] - cm.ranges.heading[0], cm.ranges.heading[0]
    # robot (lr:216-234, 843-849)
    for i, name in enumerate(robot.dof_names):
        p.default_dof_pos[i] = cfg.init_state.default_joint_angles[name]
        kp = kd = 0.0
        for key in cfg.control.stiffness:
            if key in name:
                kp, kd = cfg.control.stiffness[key], cfg.control.damping[key]
        p.p_gains[i], p.d_gains[i] = kp, kd
        p.torque_limits[i] = float(np.float32(robot.dof_effort[i])) * cfg.safety.torque_limit
        p.dof_vel_limits[i] = float(np.float32(robot.dof_velocity[i])) * cfg.safety.vel_limit
    # rewards
    scales = reward_scales(cfg, dt)
    mask = 0
    for name, s in scales.items():
        if name not in TERM_NAMES:
            raise NotImplementedError(f"reward term {name!r} is not a t1_dh_stand term")
        if name == "dof_vel_limits":
            raise NotImplementedError("_reward_dof_vel_limits reads rewards.soft_dof_vel_limit, which t1_cfg lacks")
        t = TERM_NAMES.index(name)
        p.reward_scale[t] = s
        mask |= 1 << t
    p.term_mask = mask
    p.tracking_lin_vel_scale = scales.get("tracking_lin_vel", 0.0)
    p.cmd_curriculum_max = cm.max_curriculum
    for k in ("base_height_target", "foot_min_dist", "foot_max_dist", "knee_min_dist", "knee_max_dist",
              "target_joint_pos_scale", "target_feet_height", "target_feet_height_max", "tracking_sigma",
              "max_contact_force"):
        setattr(p, k, getattr(rw, k))
    p.target_joint_pos_scale2 = 2 * rw.target_joint_pos_scale
    # observations (t1:326-357, lr:193)
    os_, ns_ = nz.obs_scales, cfg.noise.noise_scales
    p.obs_lin_vel, p.obs_ang_vel, p.obs_dof_pos, p.obs_dof_vel = os_.lin_vel, os_.ang_vel, os_.dof_pos, os_.dof_vel
    p.obs_quat, p.obs_height = os_.quat, os_.height_measurements
    for i, v in enumerate((os_.lin_vel, os_.lin_vel, os_.ang_vel)):
        p.cmd_scale[i] = v
    K, D, nc = cfg.env.num_single_obs, cfg.env.num_actions, cfg.env.num_commands
    assert K == nc + 3 * D + 6 and K <= 64
    nv = [0.0] * K
    nv[nc:nc + D] = [ns_.dof_pos * os_.dof_pos] * D
    nv[nc + D:nc + 2 * D] = [ns_.dof_vel * os_.dof_vel] * D
    nv[nc + 3 * D:nc + 3 * D + 3] = [ns_.ang_vel * os_.ang_vel] * 3
    nv[nc + 3 * D + 3:nc + 3 * D + 6] = [ns_.quat * os_.quat] * 3
    for i, v in enumerate(nv):
        p.noise_vec[i] = v
    p.noise_level = cfg.noise.noise_level
    st = cfg.init_state
    for i, v in enumerate(st.pos + st.rot + st.lin_vel + st.ang_vel):
        p.base_init_state[i] = v

    def pair(rng):
        return rng[1] - rng[0], rng[0]

    p.torque_multi_w, p.torque_multi_lo = pair(dr.torque_multiplier_range)
    p.motor_offset_w, p.motor_offset_lo = pair(dr.motor_offset_range)
    p.kp_mult_w, p.kp_mult_lo = pair(dr.stiffness_multiplier_range)
    p.kd_mult_w, p.kd_mult_lo = pair(dr.damping_multiplier_range)
    p.coulomb_w, p.coulomb_lo = pair(dr.joint_coulomb_range)
    p.viscous_w, p.viscous_lo = pair(dr.joint_viscous_range)
    for i in range(robot.num_dof):
        rng = getattr(dr, f"joint_{i + 1}_armature_range", None) or dr.joint_armature_range
        p.armature_w[i], p.armature_lo[i] = pair(rng)
    p.dof_reset_w, p.dof_reset_lo = pair((-0.1, 0.1))
    half = cfg.terrain.platform / 3 if cfg.terrain.curriculum else cfg.terrain.terrain_length / 2
    p.root_xy_w, p.root_xy_lo = pair((-half, half))
    p.push_vel_w, p.push_vel_lo = pair((-dr.max_push_vel_xy, dr.max_push_vel_xy))
    p.push_ang_w, p.push_ang_lo = pair((-dr.max_push_ang_vel, dr.max_push_ang_vel))
    if dr.add_ext_force:
        for i, (lo, hi) in enumerate(((-dr.ext_force_max_x / 2, dr.ext_force_max_x),
                                      (-dr.ext_force_max_y, dr.ext_force_max_y),
                                      (-dr.ext_force_max_z, dr.ext_force_max_z))):
            p.ext_f_w[i], p.ext_f_lo[i] = hi - lo, lo
        p.ext_t_w, p.ext_t_lo = pair((-dr.ext_torque_max, dr.ext_torque_max))
        p.ext_force_div, p.ext_torque_div = dr.ext_force_max_x + 0.1, dr.ext_torque_max + 0.1
    # terrain (lr:1574-1587, 1138-1158)
    t = cfg.terrain
    p.border_size, p.horizontal_scale, p.vertical_scale = t.border_size, t.horizontal_scale, t.vertical_scale
    if terrain is not None:
        p.terrain_env_length = terrain.env_length
        p.terrain_rows, p.terrain_cols = terrain.env_origins.shape[0], terrain.env_origins.shape[1]
        p.max_terrain_level = t.num_rows                                      # lr:1492
    return p
