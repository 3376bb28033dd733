"""Shared test plumbing: golden-fixture access, env construction in parity mode, comparisons."""
import os
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SIM_KEYS = ("root_states", "dof_state", "contact_forces", "rigid_state")
RTOL, ATOL = 1e-5, 2e-6      # north_star: <= 1e-5 relative on floats; atol covers values near zero


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"t1_step_{name}.npz"))
    steps = sorted({k.split(".")[0] for k in z.files if k.startswith("in")})
    state0 = {k[len("state0."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("state0.")}
    inputs, outputs = [], []
    for s in steps:
        inputs.append({k.split(".", 1)[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(s + ".")})
        o = "out" + s[2:]
        outputs.append({k.split(".", 1)[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(o + ".")})
    final = {k[len("final."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("final.")}
    return state0, inputs, outputs, final


def scenario_cfg(name, num_envs, frame_stack=66):
    """The product config edited like oracle/pin_against_reference.py SCENARIOS."""
    from ti5_isaacgym_b200.envs import make_t1_cfg
    cfg = make_t1_cfg(frame_stack=frame_stack)()
    cfg.env.num_envs = num_envs
    cfg.terrain.mesh_type = "trimesh" if name.startswith("trimesh") else "plane"
    if name == "plane_extra_terms":
        cfg.rewards.scales.feet_stumble = -0.5
        cfg.rewards.scales.stand_sysmetry = 0.3
        cfg.rewards.scales.termination = -1.0
    if name in ("trimesh_heights_push", "trimesh_windows"):
        cfg.terrain.measure_heights = True
        cfg.env.num_privileged_obs = 3 * (73 + 187)
        cfg.domain_rand.push_robots = True
    if name in ("plane_heights", "trimesh_no_curriculum"):
        cfg.terrain.measure_heights = True
        cfg.env.num_privileged_obs = 3 * (73 + 187)
    if name == "trimesh_no_curriculum":
        cfg.terrain.curriculum = False
    if name == "plane_windows":
        cfg.domain_rand.push_robots = True
    if name == "plane_heading":
        cfg.commands.heading_command = True
    if name == "plane_no_sw":
        cfg.commands.sw_switch = False
    if name == "plane_ref_actions":
        cfg.env.use_ref_actions = True
    if name == "plane_joint_props":
        cfg.domain_rand.randomize_joint_friction = True
        cfg.domain_rand.randomize_joint_damping = True
    if name == "plane_lag_perstep":       # _perstep
        for k in ("lag", "dof_lag", "imu_lag"):
            setattr(cfg.domain_rand, f"randomize_{k}_timesteps_perstep", True)
    if name == "plane_pos_vel_lag":       # _pos_vel_lag
        cfg.domain_rand.add_dof_lag = False
        cfg.domain_rand.add_dof_pos_vel_lag = True
        cfg.domain_rand.randomize_dof_pos_lag_timesteps_perstep = True
        cfg.domain_rand.randomize_dof_vel_lag_timesteps_perstep = False
    if name == "trimesh_points77":        # _points77
        cfg.terrain.measure_heights = True
        cfg.terrain.measured_points_x = [-0.5 + 0.1 * i for i in range(11)]
        cfg.terrain.measured_points_y = [-0.3 + 0.1 * i for i in range(7)]
        cfg.terrain.num_height = 77
        cfg.env.num_privileged_obs = 3 * (73 + 77)
        cfg.domain_rand.push_robots = True
    if name == "plane_gaits4":            # _gaits4
        cfg.commands.gait = ["walk_sagittal", "rotate", "walk_lateral", "stand"]
        cfg.commands.ranges.lin_vel_x = [-0.3, 0.8]
        cfg.commands.ranges.lin_vel_y = [-0.2, 0.2]
        cfg.commands.ranges.ang_vel_yaw = [-0.7, 0.4]
        cfg.commands.stand_com_threshold = 0.1
        cfg.commands.max_curriculum = 1.0
    if name == "plane_params":            # like oracle/pin_against_reference.py _other_params
        c = cfg
        c.control.decimation = 4
        c.control.action_scale = 0.4
        c.env.c_frame_stack = 5
        c.env.num_privileged_obs = 5 * 73
        c.env.episode_length_s = 10
        c.domain_rand.lag_timesteps_range = [1, 8]
        c.domain_rand.dof_lag_timesteps_range = [0, 5]
        c.domain_rand.imu_lag_timesteps_range = [2, 4]
        c.rewards.cycle_time = 0.64
        c.rewards.tracking_sigma = 4
        c.rewards.max_contact_force = 300
        c.normalization.clip_observations = 18.0
        c.normalization.clip_actions = 1.5
        c.noise.noise_level = 0.5
    if name == "plane_h15":
        cfg = scenario_cfg("plane_default", num_envs, frame_stack=15)
    dr = cfg.domain_rand
    if name == "plane_flags_off":         # like oracle/pin_against_reference.py _flags_off
        for f in ("add_lag", "add_dof_lag", "add_imu_lag", "randomize_gains", "randomize_coulomb_friction", "randomize_torque",
                  "randomize_motor_offset", "randomize_joint_armature", "add_ext_force"):
            setattr(dr, f, False)
        cfg.noise.add_noise = False
        cfg.rewards.only_positive_rewards = False
        cfg.commands.curriculum = False
    if name == "plane_flags_mixed":       # _flags_mixed
        for f in ("randomize_lag_timesteps", "randomize_dof_lag_timesteps", "randomize_imu_lag_timesteps", "randomize_gains",
                  "randomize_torque", "randomize_joint_armature", "add_ext_force"):
            setattr(dr, f, False)
        dr.push_robots = True
    return cfg


GOLDEN_SCENARIOS = ["plane_default", "plane_events", "trimesh_heights_push", "plane_extra_terms", "plane_windows",
                    "trimesh_windows", "plane_heading", "plane_no_sw", "plane_flags_off", "plane_flags_mixed",
                    "plane_heights", "trimesh_plain", "trimesh_no_curriculum", "plane_ref_actions", "plane_h15", "plane_params",
                    "trimesh_points77", "plane_gaits4", "plane_lag_perstep", "plane_pos_vel_lag", "plane_joint_props"]


def gym_calls_of(out):
    """The reference's gym tensor-API call names of one recorded step, in order."""
    return bytes(out["gym_calls"].numpy().tolist()).decode().split(",")


def make_env(cfg, rng_mode="pools", div_mode="ieee", **kw):
    from ti5_isaacgym_b200.envs import T1DHStandEnv
    kw.setdefault("materialize_obs", False)         # most tests read the ring views right after the step
    from ti5_isaacgym_b200.sim.synthetic import SimParams
    return T1DHStandEnv(cfg, SimParams(dt=cfg.sim.dt), 1, "cuda:0", True, rng_mode=rng_mode, div_mode=div_mode,
                        use_cuda_graph=False, **kw)


def pools_of(inp):
    return {k[4:]: v for k, v in inp.items() if k.startswith("rng_")}


def set_sim(env, inp):
    env.root_states.copy_(inp["root_states"].to(env.device))
    env.dof_state.copy_(inp["dof_state"].to(env.device))
    env.contact_forces.copy_(inp["contact_forces"].view_as(env.contact_forces).to(env.device))
    env.rigid_state.copy_(inp["rigid_state"].view_as(env.rigid_state).to(env.device))


def _same_device(a, b):
    """Compare on the GPU when either side lives there (the 65536-env windows are ~1 GB each)."""
    a, b = a.detach(), b.detach()
    dev = a.device if a.is_cuda else b.device
    return a.to(dev), b.to(dev)


def close(a, b, what, rtol=RTOL, atol=ATOL):
    a, b = _same_device(a, b)
    a, b = a.float(), b.float()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = (err > tol) | (torch.isnan(a) != torch.isnan(b))
    if bool(bad.any()):
        i = int(err.argmax())
        raise AssertionError(f"{what}: {int(bad.sum())} of {bad.numel()} beyond rtol={rtol} atol={atol}; "
                             f"worst |err|={float(err.max()):.3e} at {np.unravel_index(i, a.shape)} "
                             f"(got {float(a.flatten()[i]):.8g}, want {float(b.flatten()[i]):.8g})")


def exact(a, b, what):
    a, b = _same_device(a, b)
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.equal(a.to(b.dtype), b), f"{what}: {int((a.to(b.dtype) != b).sum())} of {b.numel()} entries differ"


def oracle_robot():
    from ti5_isaacgym_b200.envs.t1.t1_robot import robot_constants
    return robot_constants


def state_from_oracle(S, C):
    """Flat reference-named state dict of an oracle state (same keys as the golden `state0.*`)."""
    keys = ("torques actions last_actions last_last_actions last_dof_vel last_root_vel commands feet_air_time "
            "feet_height last_contacts contact_filt base_quat base_lin_vel base_ang_vel projected_gravity "
            "base_euler_xyz feet_euler_xyz ext_forces ext_torques rand_push_force rand_push_torque ref_dof_pos "
            "gait_time gait_start torque_multi motor_offsets randomized_p_gains randomized_d_gains "
            "randomized_joint_coulomb randomized_joint_viscous joint_armatures lag_buffer dof_lag_buffer "
            "imu_lag_buffer lag_timestep dof_lag_timestep imu_lag_timestep episode_length_buf phase_length_buf "
            "rew_buf reset_buf time_out_buf env_origins env_frictions body_mass").split()
    out = {k: getattr(S, k).clone() for k in keys}
    from oracle.t1_oracle import OPTIONAL_LAG_STATE          # state of the lag options t1_cfg leaves off
    out.update({k: getattr(S, k).clone() for k in OPTIONAL_LAG_STATE})
    out["ref_action"] = S.ref_action.clone()
    out["last_feet_z"] = torch.zeros(S.N, 2) if isinstance(S.last_feet_z, int) else S.last_feet_z.clone()
    out["obs_history"], out["critic_history"] = S.obs_history.clone(), S.critic_history.clone()
    out["episode_sums"] = torch.stack([S.episode_sums[k] for k in C.reward_scales], 0)
    out["counters"] = torch.tensor([S.common_step_counter, int(S.is_first_add_force), int(S.is_first_push)])
    out["command_ranges"] = torch.tensor([S.command_ranges[k] for k in ("lin_vel_x", "lin_vel_y", "ang_vel_yaw")],
                                         dtype=torch.float64)
    if hasattr(S, "terrain_levels"):
        out["terrain_levels"], out["terrain_types"] = S.terrain_levels.clone(), S.terrain_types.clone()
    return out
