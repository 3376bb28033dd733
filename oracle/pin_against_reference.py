"""Pin the oracle against the reference itself, and write the golden fixtures.

Runs ONLY in the build container (needs /root/reference).  For each scenario it steps the
unmodified reference `T1DHStandEnv` (CPU, fake gym, pooled RNG) and the oracle side by side
on identical synthetic simulator tensors and requires BIT-EQUAL results on every output and
on the persistent state after every step.  With `--write` it stores the inputs and the
reference's outputs of the small scenarios under tests/golden/ (the reference cannot travel
to the GPU box; the fixtures can).

    python oracle/pin_against_reference.py            # check only
    python oracle/pin_against_reference.py --write    # check + regenerate tests/golden/*.npz
"""
import argparse
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import t1_oracle as O                                        # noqa: E402
from oracle.reference_driver import ReferenceDriver, adopt_reference_state  # noqa: E402
from ti5_isaacgym_b200.sim.synthetic import fill_synthetic_state, synthetic_actions  # noqa: E402

STATE_KEYS = ("torques actions last_actions last_last_actions last_dof_vel last_root_vel commands feet_air_time "
              "feet_height last_contacts contact_filt base_quat base_lin_vel base_ang_vel projected_gravity "
              "base_euler_xyz feet_euler_xyz ext_forces ext_torques rand_push_force rand_push_torque ref_dof_pos ref_action "
              "gait_time gait_start torque_multi motor_offsets randomized_p_gains randomized_d_gains "
              "randomized_joint_coulomb randomized_joint_viscous joint_armatures lag_buffer dof_lag_buffer "
              "imu_lag_buffer lag_timestep dof_lag_timestep imu_lag_timestep episode_length_buf phase_length_buf "
              "rew_buf reset_buf time_out_buf env_origins").split()
# state of the options t1_cfg leaves off: compared when the reference allocated it, recorded when the option is on
OPTIONAL_KEYS = O.OPTIONAL_LAG_STATE


def robot_from_env(env):
    """Robot constants as the reference derived them from the URDF via the fake gym."""
    props = env.gym.asset.dof_props
    return SimpleNamespace(
        dof_names=list(env.dof_names), body_names=list(env.gym.asset.body_names), num_dof=env.num_dof,
        num_bodies=env.num_bodies, feet_indices=env.feet_indices.tolist(), knee_indices=env.knee_indices.tolist(),
        penalised_contact_indices=env.penalised_contact_indices.tolist(),
        termination_contact_indices=env.termination_contact_indices.tolist(),
        dof_effort=[float(x) for x in props["effort"]], dof_velocity=[float(x) for x in props["velocity"]],
        dof_lower=[float(x) for x in props["lower"]], dof_upper=[float(x) for x in props["upper"]])


def same(a, b):
    if isinstance(a, (int, float, bool)) or isinstance(b, (int, float, bool)):
        return a == b
    return a.shape == b.shape and a.dtype == b.dtype and torch.equal(a, b)


def snapshot_state(S, C):
    """All persistent state of an oracle state as a flat dict of tensors."""
    out = {k: getattr(S, k).clone() for k in STATE_KEYS}
    if C.pos_vel_lag or any(C.perstep.values()) or any(C.joint_props):
        out.update({k: getattr(S, k).clone() for k in OPTIONAL_KEYS})
    out["last_feet_z"] = torch.zeros(S.N, 2) if isinstance(S.last_feet_z, int) else S.last_feet_z.clone()
    out["obs_history"], out["critic_history"] = S.obs_history.clone(), S.critic_history.clone()
    out["episode_sums"] = torch.stack([S.episode_sums[k] for k in C.reward_scales], 0)
    out["env_frictions"], out["body_mass"] = S.env_frictions.clone(), S.body_mass.clone()
    out["counters"] = torch.tensor([S.common_step_counter, int(S.is_first_add_force), int(S.is_first_push)])
    out["command_ranges"] = torch.tensor([S.command_ranges[k] for k in ("lin_vel_x", "lin_vel_y", "ang_vel_yaw")],
                                         dtype=torch.float64)
    if hasattr(S, "terrain_levels"):
        out["terrain_levels"], out["terrain_types"] = S.terrain_levels.clone(), S.terrain_types.clone()
    return out


def _flags_off(c):
    """Every optional branch of the path switched the other way from t1_cfg."""
    dr = c.domain_rand
    for f in ("add_lag", "add_dof_lag", "add_imu_lag", "randomize_gains", "randomize_coulomb_friction", "randomize_torque",
              "randomize_motor_offset", "randomize_joint_armature", "add_ext_force"):
        setattr(dr, f, False)
    c.noise.add_noise = False
    c.rewards.only_positive_rewards = False
    c.commands.curriculum = False


def _flags_mixed(c):
    """Lags on but with fixed (maximal) indices, half of the actuator randomisation off, pushes on."""
    dr = c.domain_rand
    for f in ("randomize_lag_timesteps", "randomize_dof_lag_timesteps", "randomize_imu_lag_timesteps", "randomize_gains",
              "randomize_torque", "randomize_joint_armature", "add_ext_force"):
        setattr(dr, f, False)
    dr.push_robots = True


def _other_params(c):
    """Run-time parameters away from t1_cfg's values: decimation, critic frame stack, lag ranges, episode length, gait
    cycle, action scale, clipping, noise level, tracking sigma."""
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


def _points77(c):
    """An 11 x 7 scan grid: a 150-float privileged frame (not a multiple of four floats)."""
    c.terrain.measure_heights = True
    c.terrain.measured_points_x = [-0.5 + 0.1 * i for i in range(11)]
    c.terrain.measured_points_y = [-0.3 + 0.1 * i for i in range(7)]
    c.terrain.num_height = 77
    c.env.num_privileged_obs = 3 * (73 + 77)
    c.domain_rand.push_robots = True


def _gaits4(c):
    """Four gait slots of the kinds t1_cfg does not schedule, other command ranges and stand threshold."""
    c.commands.gait = ["walk_sagittal", "rotate", "walk_lateral", "stand"]
    c.commands.ranges.lin_vel_x = [-0.3, 0.8]
    c.commands.ranges.lin_vel_y = [-0.2, 0.2]
    c.commands.ranges.ang_vel_yaw = [-0.7, 0.4]
    c.commands.stand_com_threshold = 0.1
    c.commands.max_curriculum = 1.0


def _perstep(c):
    """Lag indices re-drawn every substep (actions) / every step (joint state, IMU): lr:1038-1043, t1:408-413, 437-442."""
    for k in ("lag", "dof_lag", "imu_lag"):
        setattr(c.domain_rand, f"randomize_{k}_timesteps_perstep", True)


def _pos_vel_lag(c):
    """Separate position / velocity lags instead of the common joint-state lag (lr:425-430, t1:416-431), re-drawn per step."""
    c.domain_rand.add_dof_lag = False
    c.domain_rand.add_dof_pos_vel_lag = True
    c.domain_rand.randomize_dof_pos_lag_timesteps_perstep = True
    c.domain_rand.randomize_dof_vel_lag_timesteps_perstep = False


SCENARIOS = {
    # name: (num_envs, steps, mesh_type, cfg edits, base-contact rate, forced events)
    "plane_default": dict(N=16, steps=28, mesh="plane"),
    "plane_events": dict(N=24, steps=40, mesh="plane", contact_rate=0.05, events=True),
    "trimesh_heights_push": dict(N=16, steps=24, mesh="trimesh", contact_rate=0.05, events=True,
                                 edit=lambda c: (setattr(c.terrain, "measure_heights", True),
                                                 setattr(c.env, "num_privileged_obs", 3 * (73 + 187)),
                                                 setattr(c.domain_rand, "push_robots", True))),
    # multi-step disturbance windows (t1:193-247) and a command-curriculum change (lr:1160-1169): at
    # common_step_counter >= 240000 / 288000 the push window lasts 20 steps (push_duration[4] / dt) and the external-force
    # window 16 (one draw step + add_duration[3] / dt apply steps); 288000 is a multiple of both intervals and of
    # max_episode_length, so both windows open and the curriculum is evaluated on the step env 0 times out
    "plane_windows": dict(N=24, steps=30, mesh="plane", contact_rate=0.05, events=True, counter=287997, track_sums=True,
                          edit=lambda c: setattr(c.domain_rand, "push_robots", True)),
    "trimesh_windows": dict(N=16, steps=26, mesh="trimesh", contact_rate=0.05, events=True, counter=287997,
                            track_sums=True,
                            edit=lambda c: (setattr(c.terrain, "measure_heights", True),
                                            setattr(c.env, "num_privileged_obs", 3 * (73 + 187)),
                                            setattr(c.domain_rand, "push_robots", True))),
    # heading mode (t1:141-176, 185-188; off in t1_cfg): the gait schedule draws a heading target, the yaw-rate command is
    # recomputed from the heading error for every env on every step
    "plane_heading": dict(N=24, steps=30, mesh="plane", contact_rate=0.05, events=True,
                          edit=lambda c: setattr(c.commands, "heading_command", True)),
    # commands.sw_switch = False (t1:89-90): the gait phase follows the episode counter, standing envs keep cycling
    "plane_no_sw": dict(N=24, steps=20, mesh="plane", contact_rate=0.05, events=True,
                        edit=lambda c: setattr(c.commands, "sw_switch", False)),
    # the optional branches of the path the other way round (no lags / actuator randomisation / noise / external force, raw
    # reward sum, fixed command ranges), and a mix (lags with fixed maximal indices, part of the randomisation off)
    "plane_flags_off": dict(N=24, steps=20, mesh="plane", contact_rate=0.05, events=True, edit=_flags_off),
    "plane_flags_mixed": dict(N=24, steps=20, mesh="plane", contact_rate=0.05, events=True, edit=_flags_mixed),
    # terrain variants: measured heights on a plane (identically zero, lr:1564), trimesh without heights, trimesh with
    # heights but without the terrain curriculum
    "plane_heights": dict(N=16, steps=12, mesh="plane", contact_rate=0.05, events=True,
                          edit=lambda c: (setattr(c.terrain, "measure_heights", True),
                                          setattr(c.env, "num_privileged_obs", 3 * (73 + 187)))),
    "trimesh_plain": dict(N=16, steps=12, mesh="trimesh", contact_rate=0.05, events=True),
    "trimesh_no_curriculum": dict(N=16, steps=12, mesh="trimesh", contact_rate=0.05, events=True,
                                  edit=lambda c: (setattr(c.terrain, "measure_heights", True),
                                                  setattr(c.env, "num_privileged_obs", 3 * (73 + 187)),
                                                  setattr(c.terrain, "curriculum", False))),
    # env.use_ref_actions (t1:360-366): the policy output is an offset on the gait's reference action
    # (switched on after reset(): the reference's own reset() steps once before `ref_action` exists and raises with it)
    "plane_ref_actions": dict(N=16, steps=14, mesh="plane", contact_rate=0.05, events=True,
                              after_reset=lambda c: setattr(c.env, "use_ref_actions", True)),
    # a shorter observation history (BASELINE config 5: frame_stack sweep)
    "plane_h15": dict(N=16, steps=20, mesh="plane", contact_rate=0.05, events=True,
                      edit=lambda c: (setattr(c.env, "frame_stack", 15), setattr(c.env, "num_observations", 15 * 47))),
    # run-time parameters away from t1_cfg's values
    "plane_params": dict(N=24, steps=24, mesh="plane", contact_rate=0.05, events=False, edit=_other_params),
    "trimesh_points77": dict(N=20, steps=14, mesh="trimesh", contact_rate=0.05, events=True, edit=_points77),
    "plane_gaits4": dict(N=24, steps=30, mesh="plane", contact_rate=0.05, events=True, edit=_gaits4),
    "plane_lag_perstep": dict(N=24, steps=24, mesh="plane", contact_rate=0.05, events=True, edit=_perstep),
    "plane_pos_vel_lag": dict(N=24, steps=24, mesh="plane", contact_rate=0.05, events=True, edit=_pos_vel_lag),
    # joint friction / damping multipliers handed to the simulator with the armatures (lr:755-773, 915-931)
    "plane_joint_props": dict(N=16, steps=14, mesh="plane", contact_rate=0.08, events=True,
                              edit=lambda c: (setattr(c.domain_rand, "randomize_joint_friction", True),
                                              setattr(c.domain_rand, "randomize_joint_damping", True))),
    "big_plane": dict(N=512, steps=12, mesh="plane", contact_rate=0.03, events=True, golden=False),
    # the reward terms the task defines but t1_cfg leaves at zero scale (t1:894-896, 917-925, 937-940)
    "plane_extra_terms": dict(N=16, steps=12, mesh="plane", contact_rate=0.08, events=True,
                              edit=lambda c: (setattr(c.rewards.scales, "feet_stumble", -0.5),
                                              setattr(c.rewards.scales, "stand_sysmetry", 0.3),
                                              setattr(c.rewards.scales, "termination", -1.0))),
}


def run_scenario(name, spec, write_dir=None, verbose=True):
    N, steps = spec["N"], spec["steps"]
    drv = ReferenceDriver(N, mesh_type=spec["mesh"], cfg_edit=spec.get("edit"), seed=11)
    env = drv.env
    terrain = None
    heights = None
    if spec["mesh"] == "trimesh":
        terrain = SimpleNamespace(env_length=env.terrain.env_length, max_level=env.max_terrain_level,
                                  origins=env.terrain_origins)
        heights = env.height_samples
    C = O.make_consts(drv.cfg, drv.cfg.sim.dt, robot_from_env(env), terrain=terrain)
    gen = torch.Generator().manual_seed(1234)
    # the reference's own reset() (construction-time state is already random)
    fill_synthetic_state(drv.sim, env.env_origins, gen)
    pools = O.draw_pools(C, N, gen)
    with drv.pooled_rng(pools):
        env.reset()
    # spread episode phases so that stand phases, gait switches and time-outs all occur (8d)
    env.episode_length_buf[:] = torch.randint(1, 2000, (N,), generator=gen)
    if spec.get("events"):
        env.episode_length_buf[0] = 2398                  # time-out at the 3rd step
        env.episode_length_buf[1] = int(env.gait_time[1, 1]) - 2       # gait switch to "stand"
        env.episode_length_buf[2] = int(env.gait_time[2, 2]) - 3       # and back to walking
        if env.gait_time.shape[1] > 3:                                 # a fourth gait slot
            env.episode_length_buf[3] = int(env.gait_time[3, 3]) - 4
            env.episode_length_buf[4] = int(env.gait_time[4, 1]) - 5
        env.phase_length_buf[:] = env.episode_length_buf
        env.common_step_counter = spec.get("counter", 2397)   # command-curriculum check at step 3, ext-force window
    if spec.get("track_sums"):        # tracking reward above 80 % of its maximum: the curriculum widens lin_vel_x
        env.episode_sums["tracking_lin_vel"][:] = 0.9 * env.reward_scales["tracking_lin_vel"] * float(env.max_episode_length)
    if spec.get("after_reset"):
        spec["after_reset"](env.cfg)
    env.gym.log_calls = True
    S = adopt_reference_state(O.new_state(C, N), env)
    state0 = snapshot_state(S, C)
    if terrain is not None:
        state0["terrain_origins"] = terrain.origins.clone()
        state0["terrain_env_length"] = torch.tensor(float(terrain.env_length))
    rec = dict(inputs=[], outputs=[])
    cov = dict(time_outs=0, stand_env_steps=0, ext_force_steps=0, ext_apply_steps=0, push_steps=0, curriculum_changes=0, gait_switches=0)
    for t in range(steps):
        fill_synthetic_state(drv.sim, env.env_origins, gen, base_contact_rate=spec.get("contact_rate", 0.01))
        sim0 = {k: getattr(drv.sim, k).clone() for k in ("root_states", "dof_state", "contact_forces", "rigid_state")}
        actions = synthetic_actions(N, gen, "cpu")
        pools = O.draw_pools(C, N, gen)
        # oracle on a private copy of the simulator tensors
        osim = SimpleNamespace(**{k: v.clone() for k, v in sim0.items()})
        o_obs, o_priv, o_rew, o_reset, o_extras = O.step(C, S, osim, actions.clone(), pools, terrain=terrain, height_samples=heights)
        env.gym.calls.clear()
        ranges_before = {k: list(v) for k, v in env.command_ranges.items()}
        r_obs, r_priv, r_rew, r_reset, r_extras = drv.step(actions, pools)
        calls = list(env.gym.calls)
        bad = []
        # what the reference handed to apply_rigid_body_force_tensors (t1:247): body 0 rows == the oracle's applied_*
        forces = [c[1] for c in calls if c[0] == "apply_rigid_body_force_tensors"]
        r_af, r_at = torch.zeros(N, 3), torch.zeros(N, 3)
        if forces:
            assert len(forces) == 1 and float(forces[0][0][:, 1:].abs().sum()) == 0 and float(forces[0][1][:, 1:].abs().sum()) == 0
            r_af, r_at = forces[0][0][:, 0].clone(), forces[0][1][:, 0].clone()
        if not (same(S.applied_force, r_af) and same(S.applied_torque, r_at)):
            bad.append("applied_force/torque")
        pushed = [c[1] for c in calls if c[0] == "set_actor_root_state_tensor"]
        # lr:915-939: the per-env property structs handed back to the simulator for the re-spawned envs
        props = [c[1] for c in calls if c[0] == "set_actor_dof_properties"]
        props_env = torch.tensor([e for e, _ in props], dtype=torch.int64)
        props_arm = torch.tensor(np.stack([d["armature"] for _, d in props]) if props else np.zeros((0, 12), np.float32))
        if not (same(props_env, S.reset_ids) and same(props_arm, S.joint_armatures[S.reset_ids])):
            bad.append("dof props")
        # lr:921-931: friction / damping of the asset times the env's multiplier (the fake gym hands out the asset's values)
        asset = env.gym.asset.dof_props
        for fld, on, coeff in (("friction", C.joint_props[0], S.joint_friction_coeffs), ("damping", C.joint_props[1], S.joint_damping_coeffs)):
            got = torch.tensor(np.stack([d[fld] for _, d in props]) if props else np.zeros((0, 12), np.float32))
            want = torch.from_numpy(asset[fld].copy()).unsqueeze(0) * (coeff[S.reset_ids] if on else torch.ones(len(S.reset_ids), 1))
            if not same(got, want.float()):
                bad.append("dof props " + fld)
        for nm in ("set_dof_state_tensor_indexed", "set_actor_root_state_tensor_indexed"):
            for c in calls:
                if c[0] == nm and not (same(c[1][1].long(), S.reset_ids) and c[1][2] == len(S.reset_ids)):
                    bad.append(nm + " ids")
        for key, a, b in (("obs", o_obs, r_obs), ("priv", o_priv, r_priv), ("rew", o_rew, r_rew), ("reset", o_reset, r_reset)):
            if not same(a, b):
                bad.append(key)
        for key in list(STATE_KEYS) + [k for k in OPTIONAL_KEYS if hasattr(env, k)]:
            if hasattr(env, key) and not same(getattr(S, key), getattr(env, key)):
                bad.append(key)
        if not same(S.last_feet_z, env.last_feet_z):
            bad.append("last_feet_z")
        for key in env.episode_sums:
            if not same(S.episode_sums[key], env.episode_sums[key]):
                bad.append("episode_sums." + key)
        for key in ("root_states", "dof_state"):
            if not same(getattr(osim, key), getattr(drv.sim, key)):
                bad.append("sim." + key)
        if not same(torch.stack(list(env.obs_history), 0), S.obs_history):
            bad.append("obs_history")
        if not same(torch.stack(list(env.critic_history), 0), S.critic_history):
            bad.append("critic_history")
        if "episode" in r_extras:
            for k, v in r_extras["episode"].items():
                ov = o_extras["episode"][k]
                if not (same(v, ov) if torch.is_tensor(v) else v == ov):
                    bad.append("extras." + k)
            if not same(r_extras["time_outs"], o_extras["time_outs"]):
                bad.append("extras.time_outs")
        if S.command_ranges != {k: list(v) for k, v in env.command_ranges.items()}:
            bad.append("command_ranges")
        if not isinstance(S.measured_heights, int) and not same(S.measured_heights, env.measured_heights):
            bad.append("measured_heights")
        cov["time_outs"] += int(env.time_out_buf.sum())
        cov["stand_env_steps"] += int(O.stand_command(C, S).sum())
        cov["ext_force_steps"] += int(bool(env.ext_forces.abs().sum() > 0))
        cov["ext_apply_steps"] += int(bool(r_af.abs().sum() > 0))
        cov["push_steps"] += int(len(pushed) > 0)
        cov["curriculum_changes"] += int({k: list(v) for k, v in env.command_ranges.items()} != ranges_before)
        cov["gait_switches"] += int(((env.episode_length_buf.unsqueeze(1) == env.gait_time[:, 1:]).any(1)).sum())
        if bad:
            raise SystemExit(f"[{name}] step {t}: oracle != reference on {bad}")
        rec["inputs"].append(dict(sim0, actions=actions, **{"rng_" + k: v for k, v in pools.items()}))
        rec["outputs"].append(dict(
            obs_new=r_obs[:, -C.cfg.env.num_single_obs:].clone(), priv_new=r_priv[:, -S.critic_history.shape[2]:].clone(),
            rew=r_rew.clone(), reset=r_reset.clone(), time_out=env.time_out_buf.clone(), torques=env.torques.clone(),
            commands=env.commands.clone(), contact_filt=env.contact_filt.clone(), feet_air_time=env.feet_air_time.clone(),
            ref_dof_pos=env.ref_dof_pos.clone(), root_after=drv.sim.root_states.clone(), dof_after=drv.sim.dof_state.clone(),
            episode_sums=torch.stack([env.episode_sums[k] for k in C.reward_scales], 0),
            reward_terms=torch.stack([S.reward_terms[k] for k in C.reward_scales], 0),
            n_reset=torch.tensor(int(r_reset.sum())),
            applied_force=r_af, applied_torque=r_at, ext_forces=env.ext_forces.clone(), ext_torques=env.ext_torques.clone(),
            rand_push_force=env.rand_push_force.clone(), rand_push_torque=env.rand_push_torque.clone(),
            command_ranges=torch.tensor([env.command_ranges[k] for k in ("lin_vel_x", "lin_vel_y", "ang_vel_yaw")],
                                        dtype=torch.float64),
            props_env=props_env, props_armature=props_arm,
            props_friction_coeff=S.joint_friction_coeffs[S.reset_ids].clone(), props_damping_coeff=S.joint_damping_coeffs[S.reset_ids].clone(),
            # the gym tensor-API calls of this step, in order (lower boundary), as one string
            gym_calls=torch.tensor(list(",".join(c[0] for c in calls).encode()), dtype=torch.uint8)))
    n_resets = sum(int(o["n_reset"]) for o in rec["outputs"])
    if verbose:
        print(f"[{name}] N={N} steps={steps}: oracle == reference bit-for-bit "
              f"(resets={n_resets}, coverage={cov}, obs checksum={float(r_obs.double().sum()):.6f})")
    if write_dir is not None and spec.get("golden", True):
        flat = {f"state0.{k}": v.numpy() for k, v in state0.items()}
        for t, (i, o) in enumerate(zip(rec["inputs"], rec["outputs"])):
            for k, v in i.items():
                flat[f"in{t:03d}.{k}"] = v.numpy()
            for k, v in o.items():
                flat[f"out{t:03d}.{k}"] = v.numpy()
        flat["final.obs"] = r_obs.numpy()
        flat["final.priv"] = r_priv.numpy()
        np.savez_compressed(os.path.join(write_dir, f"t1_step_{name}.npz"), **flat)
    return rec


def pin_gae(write_dir=None):
    """GAE: `RolloutStorage.compute_returns` (rollout_storage.py:97-119) vs the oracle."""
    from oracle.reference_driver import import_reference
    import_reference()
    from humanoid.algo.ppo.rollout_storage import RolloutStorage
    g = torch.Generator().manual_seed(99)
    for T, N in ((24, 64), (24, 4096), (5, 7)):
        st = RolloutStorage(N, T, [4], [4], [2])
        st.rewards[:] = torch.randn(T, N, 1, generator=g)
        st.values[:] = torch.randn(T, N, 1, generator=g)
        st.dones[:] = (torch.rand(T, N, 1, generator=g) < 0.02).byte()
        last = torch.randn(N, 1, generator=g)
        st.compute_returns(last, 0.994, 0.9)
        ret, adv = O.gae_returns(st.rewards, st.values, st.dones, last, 0.994, 0.9)
        if not (torch.equal(ret, st.returns) and torch.equal(adv, st.advantages)):
            raise SystemExit(f"GAE oracle != reference at T={T} N={N}")
        if write_dir is not None and N == 64:
            np.savez_compressed(os.path.join(write_dir, "gae_T24_N64.npz"), rewards=st.rewards.numpy(),
                                values=st.values.numpy(), dones=st.dones.numpy(), last_values=last.numpy(),
                                returns=st.returns.numpy(), advantages=st.advantages.numpy(),
                                gamma=np.float64(0.994), lam=np.float64(0.9))
    print("[gae] oracle == reference bit-for-bit")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", action="store_true")
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    out = os.path.join(ROOT, "tests", "golden") if a.write else None
    for nm, sp in SCENARIOS.items():
        if a.only is None or a.only == nm:
            run_scenario(nm, sp, out)
    if a.only is None:
        pin_gae(out)
