"""CUDA path vs the oracle on seeded inputs, larger batches and longer horizons than the fixtures,
in both rounding modes (oracle on the CPU with IEEE division; oracle on the GPU with torch's
reciprocal-multiply), plus size-independent properties at the BASELINE size (8192 envs)."""
from types import SimpleNamespace

import pytest
import torch

from helpers import SIM_KEYS, close, exact, make_env, scenario_cfg, state_from_oracle

from oracle import t1_oracle as O

pytestmark = pytest.mark.gpu


def _build(name, N, device, seed=3, frame_stack=66, counter=2397):
    from ti5_isaacgym_b200.envs.t1.t1_robot import robot_constants
    from ti5_isaacgym_b200.sim.synthetic import SyntheticTerrain, synthetic_height_field
    cfg = scenario_cfg(name, N, frame_stack)
    terrain = heights = None
    if cfg.terrain.mesh_type == "trimesh":
        st = SyntheticTerrain(cfg.terrain, N)
        terrain = SimpleNamespace(env_length=st.env_length, max_level=cfg.terrain.num_rows,
                                  origins=torch.from_numpy(st.env_origins).float().to(device))
        heights = synthetic_height_field(st.tot_rows, st.tot_cols, seed=7).to(device)
    C = O.make_consts(cfg, cfg.sim.dt, robot_constants(cfg), device=device, terrain=terrain)
    S = O.new_state(C, N)
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=gen).to(device)
    # a plausible mid-training state: random actuator parameters, lags, gait schedule, episode phases
    S.randomized_p_gains[:] = C.p_gains * (0.8 + 0.4 * r(N, 12))
    S.randomized_d_gains[:] = C.d_gains * (0.8 + 0.4 * r(N, 12))
    S.motor_offsets[:] = -0.035 + 0.07 * r(N, 12)
    S.randomized_joint_coulomb[:] = 0.1 + 0.9 * r(N, 12)
    S.randomized_joint_viscous[:] = 0.1 + 0.8 * r(N, 12)
    S.lag_timestep[:] = torch.randint(0, 31, (N,), generator=gen).to(device)
    S.dof_lag_timestep[:] = torch.randint(0, 31, (N,), generator=gen).to(device)
    S.imu_lag_timestep[:] = torch.randint(0, 11, (N,), generator=gen).to(device)
    S.gait_start[:] = torch.randint(0, 2, (N,), generator=gen).to(device) * 0.5
    S.gait_time[:, 1] = torch.randint(700, 1100, (N,), generator=gen).to(device).int()
    S.gait_time[:, 2] = S.gait_time[:, 1] + torch.randint(200, 600, (N,), generator=gen).to(device).int()
    S.episode_length_buf[:] = torch.randint(1, 2395, (N,), generator=gen).to(device)
    S.episode_length_buf[: N // 64] = 2398                       # time-outs two steps in
    S.episode_length_buf[N // 64: N // 32] = S.gait_time[N // 64: N // 32, 1].long() - 2      # walk -> stand
    S.phase_length_buf[:] = S.episode_length_buf
    S.commands[:, :3] = -0.5 + r(N, 3)
    S.commands[S.episode_length_buf > S.gait_time[:, 1], :3] = 0
    S.commands[S.episode_length_buf > S.gait_time[:, 2], :3] = -0.5 + r(int((S.episode_length_buf > S.gait_time[:, 2]).sum()), 3)
    if cfg.commands.heading_command:
        S.commands[:, 3] = -3.0 + 6.0 * r(N)                      # heading targets of a schedule in progress
    if C.pos_vel_lag:
        S.dof_pos_lag_timestep[:] = torch.randint(7, 26, (N,), generator=gen).to(device)
        S.dof_vel_lag_timestep[:] = torch.randint(7, 26, (N,), generator=gen).to(device)
    for kind in ("lag", "dof_lag", "imu_lag", "dof_pos_lag", "dof_vel_lag"):     # a per-step re-draw in progress
        if C.perstep[kind]:
            getattr(S, f"last_{kind}_timestep")[:] = getattr(S, f"{kind}_timestep")
    S.env_frictions[:] = 0.2 + 1.1 * r(N, 1)
    S.body_mass[:] = 10 + 5 * r(N, 1)
    S.common_step_counter = counter                               # curriculum check + ext-force window at step 3
    if counter > 2397:      # "windows" scenarios: tracking sums above 80 % of the maximum, the command curriculum fires
        S.episode_sums["tracking_lin_vel"][:] = 0.9 * C.reward_scales["tracking_lin_vel"] * 2400
    if terrain is not None:
        S.terrain_levels[:] = torch.randint(0, 6, (N,), generator=gen).to(device)
        S.terrain_types[:] = torch.div(torch.arange(N), N / cfg.terrain.num_cols, rounding_mode="floor").long().to(device)
        S.env_origins[:] = terrain.origins[S.terrain_levels, S.terrain_types]
    return cfg, C, S, terrain, heights, gen


# base-contact (termination) rate per step: None = 4 % every step; a tuple is cycled — 1.0 = every env re-spawns in the
# same step (mass reset inside a step), 0.0 = a step without any reset (extras keep the previous snapshot)
MASS = (1.0, 0.0, 0.04, 0.0, 1.0, 1.0)


@pytest.mark.parametrize("name,N,steps,where,H,rates", [
    ("plane_events", 1024, 40, "cpu", 66, None), ("plane_events", 1024, 40, "cuda", 66, None),
    ("trimesh_heights_push", 512, 24, "cpu", 66, None), ("trimesh_heights_push", 512, 24, "cuda", 66, None),
    ("plane_events", 8192, 6, "cuda", 66, None),
    ("plane_events", 96, 400, "cuda", 66, None),  # long horizon: 400 consecutive steps (4000 substeps) with resets, no re-sync
    ("plane_events", 100, 12, "cpu", 5, None), ("plane_events", 333, 12, "cuda", 15, None),
    ("plane_events", 64, 8, "cuda", 100, None),   # BASELINE config 5: H sweep, ragged N
    # edge cases: a single env, tiles that are never full, mass resets and reset-free steps
    ("plane_events", 1, 12, "cuda", 66, MASS), ("plane_events", 31, 12, "cuda", 66, MASS),
    ("plane_events", 33, 12, "cuda", 66, MASS), ("plane_events", 2048, 12, "cuda", 66, MASS),
    ("trimesh_heights_push", 200, 12, "cuda", 66, MASS),
    # multi-step push / external-force windows (20 / 16 steps at common_step_counter >= 288000) with the apply branch,
    # window exit, and a command-curriculum change; in both rounding modes
    ("plane_windows", 1024, 30, "cuda", 66, None), ("plane_windows", 256, 30, "cpu", 66, None),
    ("trimesh_windows", 512, 26, "cuda", 66, None),
    # heading mode (commands.heading_command, off in t1_cfg): heading target drawn by the schedule, yaw rate from the
    # heading error on every step; in both rounding modes
    ("plane_heading", 1024, 40, "cuda", 66, None), ("plane_heading", 256, 20, "cpu", 66, None),
    ("plane_no_sw", 1024, 20, "cuda", 66, None),      # commands.sw_switch = False: phase from the episode counter
    # the optional branches the other way round from t1_cfg (no lags / randomisation / noise / force; fixed lag indices)
    ("plane_flags_off", 1024, 20, "cuda", 66, None), ("plane_flags_mixed", 1024, 20, "cuda", 66, None),
    ("plane_flags_off", 200, 12, "cpu", 66, MASS),
    # the lag options t1_cfg marks "always False": per-step re-draws of the lag indices, separate position / velocity lags
    ("plane_lag_perstep", 1024, 30, "cuda", 66, None), ("plane_lag_perstep", 200, 12, "cpu", 66, MASS),
    ("plane_pos_vel_lag", 1024, 30, "cuda", 66, None), ("plane_pos_vel_lag", 200, 12, "cpu", 66, MASS),
    # BASELINE config 3 at its own size, and the plane step at the sizes where env_block, early mode, the carve-out and
    # the 128-register build switch
    ("trimesh_heights_push", 8192, 5, "cuda", 66, None),
    ("trimesh_heights_push", 12288, 4, "cuda", 66, None),    # env_block 64 with the heights staged: four roles per env
    ("plane_events", 16384, 4, "cuda", 66, None), ("plane_events", 65536, 3, "cuda", 66, None),
])
def test_env_follows_oracle(name, N, steps, where, H, rates, fused=True):
    from ti5_isaacgym_b200.sim.synthetic import alloc_sim_tensors, fill_synthetic_state, synthetic_actions
    device = "cuda:0" if where == "cuda" else "cpu"
    cfg, C, S, terrain, heights, gen = _build(name, N, device, frame_stack=H, counter=287997 if "windows" in name else 2397)
    env = make_env(scenario_cfg(name, N, H), div_mode="ieee" if where == "cpu" else "reciprocal", fused_step=fused)
    state = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in state_from_oracle(S, C).items()}
    if terrain is not None:
        state["terrain_origins"] = terrain.origins.cpu()
    env.load_state(state)
    sim_cpu = alloc_sim_tensors(N, "cpu")
    n_resets = n_stand = n_apply = n_push = 0
    range0 = list(S.command_ranges["lin_vel_x"])
    for t in range(steps):
        fill_synthetic_state(sim_cpu, S.env_origins.cpu(), gen, base_contact_rate=0.04 if rates is None else rates[t % len(rates)])
        actions = synthetic_actions(N, gen, "cpu")
        pools = O.draw_pools(C, N, gen)
        sim = SimpleNamespace(**{k: getattr(sim_cpu, k).clone().to(device) for k in SIM_KEYS})
        for k in SIM_KEYS:
            getattr(env.gym.tensors, k).copy_(getattr(sim_cpu, k))
        env.set_rng_pools(pools)
        obs, priv, rew, reset, extras = env.step(actions.cuda())
        pools_d = {k: v.to(device) for k, v in pools.items()}
        o_obs, o_priv, o_rew, o_reset, o_extras = O.step(C, S, sim, actions.to(device), pools_d, terrain=terrain,
                                                         height_samples=heights)
        tag = f"{name}[{where}] N={N} step {t}: "
        exact(reset, o_reset, tag + "reset_buf")
        exact(env.time_out_buf, S.time_out_buf, tag + "time_out_buf")
        ids = o_reset.nonzero().flatten()
        g = env.sync_from_device()
        assert g.n_reset == len(ids)
        exact(env.reset_ids[:len(ids)], ids.to(torch.int32), tag + "reset ids")
        exact(env.contact_filt, S.contact_filt, tag + "contact_filt")
        exact(env.last_contacts, S.last_contacts, tag + "last_contacts")
        exact(env.episode_length_buf, S.episode_length_buf, tag + "episode_length_buf")
        exact(env.phase_length_buf, S.phase_length_buf, tag + "phase_length_buf")
        exact(env.gait_time, S.gait_time, tag + "gait_time")
        close(env.torques, S.torques, tag + "torques")
        close(rew, o_rew, tag + "rew_buf")
        close(obs, o_obs, tag + "obs_buf")
        close(priv, o_priv, tag + "privileged_obs_buf")
        close(env.commands, S.commands, tag + "commands")
        for nm in env.reward_scales:
            close(env.reward_terms[nm], S.reward_terms[nm], tag + "term " + nm)
            close(env.episode_sums[nm], S.episode_sums[nm], tag + "episode_sums " + nm)
        for attr in ("base_lin_vel", "base_ang_vel", "projected_gravity", "base_euler_xyz", "feet_euler_xyz", "feet_air_time",
                     "feet_height", "ref_dof_pos", "last_actions", "last_last_actions", "last_dof_vel", "last_root_vel",
                     "ext_forces", "ext_torques", "applied_force", "applied_torque", "rand_push_force", "rand_push_torque", "motor_offsets", "randomized_p_gains", "randomized_joint_viscous", "joint_armatures",
                     "gait_start", "env_origins"):
            close(getattr(env, attr), getattr(S, attr), tag + attr)
        exact(env.lag_timestep, S.lag_timestep, tag + "lag_timestep")
        close(env.lag_buffer, S.lag_buffer, tag + "lag_buffer (ring vs shifted array)")
        if C.pos_vel_lag:       # separate position / velocity lags: both live in the DOF ring
            close(env.dof_pos_lag_buffer, S.dof_pos_lag_buffer, tag + "dof_pos_lag_buffer")
            close(env.dof_vel_lag_buffer, S.dof_vel_lag_buffer, tag + "dof_vel_lag_buffer")
            exact(env.dof_pos_lag_timestep, S.dof_pos_lag_timestep, tag + "dof_pos_lag_timestep")
            exact(env.dof_vel_lag_timestep, S.dof_vel_lag_timestep, tag + "dof_vel_lag_timestep")
        else:
            close(env.dof_lag_buffer, S.dof_lag_buffer, tag + "dof_lag_buffer")
            exact(env.dof_lag_timestep, S.dof_lag_timestep, tag + "dof_lag_timestep")
        exact(env.imu_lag_timestep, S.imu_lag_timestep, tag + "imu_lag_timestep")
        if any(C.perstep.values()):
            last = env.last_lag_timesteps()
            for col, kind in enumerate(("lag", "dof_lag", "imu_lag", "dof_pos_lag", "dof_vel_lag")):
                if C.perstep[kind] and (col < 3 or C.pos_vel_lag):
                    exact(last[:, col], getattr(S, f"last_{kind}_timestep"), tag + f"last_{kind}_timestep")
        close(env.imu_lag_buffer, S.imu_lag_buffer, tag + "imu_lag_buffer")
        close(env.root_states, sim.root_states, tag + "root_states")
        close(env.dof_state, sim.dof_state, tag + "dof_state")
        if terrain is not None:
            exact(env.terrain_levels, S.terrain_levels, tag + "terrain_levels")
            close(env.measured_heights, S.measured_heights, tag + "measured_heights")
        if len(ids):
            for k, v in o_extras["episode"].items():
                got = extras["episode"][k]
                close(torch.as_tensor(got).reshape(()), torch.as_tensor(v).float().reshape(()), tag + "extras " + k)
            exact(extras["time_outs"], o_extras["time_outs"], tag + "extras time_outs")
        for k in ("lin_vel_x", "lin_vel_y", "ang_vel_yaw"):      # lr:1160-1169, read back by sync_from_device above
            assert [float(x) for x in env.command_ranges[k]] == [float(x) for x in S.command_ranges[k]], tag + "command range " + k
        n_apply += int(bool(S.applied_force.abs().sum() > 0))
        n_push += int(bool(S.rand_push_torque.abs().sum() > 0))
        n_resets += len(ids)
        n_stand += int(O.stand_command(C, S).sum())
    assert n_resets > 0 and (n_stand > 0 or rates is not None), "the scenario must exercise resets and the stand phase"
    if "windows" in name:
        assert n_apply >= 5 and n_push >= 5, "the multi-step apply / push branches must run"
        assert list(S.command_ranges["lin_vel_x"]) != range0, "the command curriculum must change the range"


@pytest.mark.parametrize("name,N,steps", [("plane_events", 1024, 20), ("trimesh_windows", 200, 26), ("plane_events", 65536, 3),
                                          ("plane_heading", 512, 12), ("plane_no_sw", 512, 12),
                                          ("plane_flags_off", 512, 12), ("plane_flags_mixed", 512, 12)])
def test_twelve_launch_sequence_follows_oracle(name, N, steps):
    """The unfused kernels of a step without a simulator (ti5_first_substep, ti5_substep, ti5_post_physics) against the
    oracle: the default path of the tests above is ti5_fused_step."""
    test_env_follows_oracle(name, N, steps, "cuda", 66, None, fused=False)


def test_command_curriculum_fires_on_device():
    """lr:1160-1169: with high tracking sums on a step where common_step_counter % 2400 == 0 the x range widens."""
    name, N = "plane_events", 256
    cfg, C, S, _, _, gen = _build(name, N, "cpu")
    S.episode_sums["tracking_lin_vel"][:] = 0.9 * C.reward_scales["tracking_lin_vel"] * 2400
    S.common_step_counter = 2399
    S.episode_length_buf[:8] = 2400
    env = make_env(scenario_cfg(name, N))
    env.load_state(state_from_oracle(S, C))
    from ti5_isaacgym_b200.sim.synthetic import alloc_sim_tensors, fill_synthetic_state, synthetic_actions
    sim = alloc_sim_tensors(N, "cpu")
    fill_synthetic_state(sim, S.env_origins, gen)
    actions, pools = synthetic_actions(N, gen, "cpu"), O.draw_pools(C, N, gen)
    for k in SIM_KEYS:
        getattr(env.gym.tensors, k).copy_(getattr(sim, k))
    env.set_rng_pools(pools)
    env.step(actions.cuda())
    O.step(C, S, sim, actions, pools)
    env.sync_from_device()
    assert S.command_ranges["lin_vel_x"] == [-0.75, 1.0]
    assert env.command_ranges["lin_vel_x"] == [-0.75, 1.0]
    close(env.commands, S.commands, "commands drawn from the widened range")


def test_simulator_in_the_loop_uses_the_unfused_entry_points():
    """With a simulator between torque and lag push (lr:401-434) the env drives ti5_begin_step /
    ti5_torque_substep / ti5_lag_push per substep; the state changes at every substep, like PhysX would."""
    from ti5_isaacgym_b200.sim.synthetic import alloc_sim_tensors, fill_synthetic_state, synthetic_actions
    name, N, steps = "plane_events", 512, 14
    cfg, C, S, _, _, gen = _build(name, N, "cpu")
    env = make_env(scenario_cfg(name, N))
    env.load_state(state_from_oracle(S, C))
    sim = alloc_sim_tensors(N, "cpu")
    for t in range(steps):
        fill_synthetic_state(sim, S.env_origins, gen, base_contact_rate=0.04)
        actions, pools = synthetic_actions(N, gen, "cpu"), O.draw_pools(C, N, gen)
        # ten successor states of the joints and the base, one per simulator substep
        dofs = [sim.dof_state + 0.01 * torch.randn(sim.dof_state.shape, generator=gen) for _ in range(10)]
        roots = []
        for _ in range(10):
            r = sim.root_states.clone()
            r[:, 3:7] += 0.01 * torch.randn(N, 4, generator=gen)
            r[:, 3:7] /= r[:, 3:7].norm(dim=1, keepdim=True)
            r[:, 7:13] += 0.02 * torch.randn(N, 6, generator=gen)
            roots.append(r)
        for k in SIM_KEYS:
            getattr(env.gym.tensors, k).copy_(getattr(sim, k))
        osim = SimpleNamespace(**{k: getattr(sim, k).clone() for k in SIM_KEYS})

        def physics_oracle(k):
            osim.dof_state.copy_(dofs[k])
            osim.root_states.copy_(roots[k])

        def physics_env(k):
            env.gym.tensors.dof_state.copy_(dofs[k])
            env.gym.tensors.root_states.copy_(roots[k])

        env.gym.physics = physics_env
        env.set_rng_pools(pools)
        obs, priv, rew, reset, _ = env.step(actions.cuda())
        o_obs, o_priv, o_rew, o_reset, _ = O.step(C, S, osim, actions, pools, physics=physics_oracle)
        tag = f"simulator in the loop, step {t}: "
        exact(reset, o_reset, tag + "reset_buf")
        close(env.torques, S.torques, tag + "torques")
        close(rew, o_rew, tag + "rew_buf")
        close(obs, o_obs, tag + "obs_buf")
        close(priv, o_priv, tag + "privileged_obs_buf")
        close(env.dof_lag_buffer, S.dof_lag_buffer, tag + "dof_lag_buffer")
        close(env.imu_lag_buffer, S.imu_lag_buffer, tag + "imu_lag_buffer")
        close(env.lag_buffer, S.lag_buffer, tag + "lag_buffer")
    env.gym.physics = None
