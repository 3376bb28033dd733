"""CUDA path vs the golden fixtures recorded from the UNMODIFIED reference (CPU torch) by
oracle/pin_against_reference.py: same initial state, same simulator tensors, same random draws.
Bit-exact on reset flags / time-outs / reset id lists / contact masks; <= 1e-5 relative on floats."""
import pytest
import torch

from helpers import GOLDEN_SCENARIOS, close, exact, load_golden, make_env, pools_of, scenario_cfg, set_sim

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fused", [True, False], ids=["fused_step", "twelve_launches"])
@pytest.mark.parametrize("name", GOLDEN_SCENARIOS)
def test_step_matches_reference_fixture(name, fused):
    state0, inputs, outputs, final = load_golden(name)
    N = state0["commands"].shape[0]
    env = make_env(scenario_cfg(name, N), fused_step=fused)
    env.load_state(state0)
    K = env.cfg.env.num_single_obs
    P = env._params.priv_frame
    for t, (inp, out) in enumerate(zip(inputs, outputs)):
        set_sim(env, inp)
        env.set_rng_pools(pools_of(inp))
        obs, priv, rew, reset, extras = env.step(inp["actions"].cuda())
        tag = f"{name} step {t}: "
        exact(reset, out["reset"], tag + "reset_buf")
        exact(env.time_out_buf, out["time_out"], tag + "time_out_buf")
        ids = out["reset"].nonzero().flatten()
        g = env.sync_from_device()
        assert g.n_reset == len(ids), tag + f"n_reset {g.n_reset} vs {len(ids)}"
        exact(env.reset_ids[:len(ids)], ids.to(torch.int32), tag + "reset ids (ascending)")
        exact(env.contact_filt, out["contact_filt"], tag + "contact_filt")
        close(env.torques, out["torques"], tag + "torques")
        close(rew, out["rew"], tag + "rew_buf")
        close(env.commands, out["commands"], tag + "commands")
        close(env.feet_air_time, out["feet_air_time"], tag + "feet_air_time")
        close(env.ref_dof_pos, out["ref_dof_pos"], tag + "ref_dof_pos")
        close(torch.stack([env.reward_terms[n] for n in env.reward_scales]), out["reward_terms"], tag + "reward terms")
        close(torch.stack([env.episode_sums[n] for n in env.reward_scales]), out["episode_sums"], tag + "episode_sums")
        close(obs[:, -K:], out["obs_new"], tag + "newest obs frame")
        close(priv[:, -P:], out["priv_new"], tag + "newest privileged frame")
        close(env.root_states, out["root_after"], tag + "root_states after resets")
        close(env.dof_state, out["dof_after"], tag + "dof_state after resets")
        # disturbance windows (t1:193-247): draws, what is handed to apply_rigid_body_force_tensors, pushes
        for k in ("applied_force", "applied_torque", "ext_forces", "ext_torques", "rand_push_force", "rand_push_torque"):
            close(getattr(env, k), out[k], tag + k)
        # command curriculum (lr:1160-1169): the ranges the NEXT step draws from
        got = torch.tensor([env.command_ranges[k] for k in ("lin_vel_x", "lin_vel_y", "ang_vel_yaw")], dtype=torch.float64)
        exact(got, out["command_ranges"], tag + "command ranges")
        assert obs.shape == (N, env.cfg.env.frame_stack * K) and priv.shape == (N, env.cfg.env.c_frame_stack * P)
    close(obs, final["obs"], name + ": full observation history (layout oldest -> newest)")
    close(priv, final["priv"], name + ": full privileged frame stack")
