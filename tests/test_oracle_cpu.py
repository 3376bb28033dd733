"""CPU tests of the oracle: it replays the golden fixtures recorded from the unmodified reference
(so the pinning can be re-checked anywhere, without /root/reference), and — where the reference tree
is mounted — it is stepped side by side with the reference itself."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from helpers import GOLDEN, GOLDEN_SCENARIOS, SIM_KEYS, load_golden, pools_of, scenario_cfg

from oracle import t1_oracle as O


def _terrain_of(state0, cfg):
    if "terrain_origins" not in state0:
        return None, None
    from ti5_isaacgym_b200.sim.synthetic import synthetic_height_field
    t = SimpleNamespace(env_length=float(state0["terrain_env_length"]), max_level=cfg.terrain.num_rows,
                        origins=state0["terrain_origins"])
    return t, synthetic_height_field(2100, 2100, seed=7)


@pytest.mark.parametrize("name", GOLDEN_SCENARIOS)
def test_oracle_replays_reference_fixture(name):
    from ti5_isaacgym_b200.envs.t1.t1_robot import robot_constants
    state0, inputs, outputs, final = load_golden(name)
    N = state0["commands"].shape[0]
    cfg = scenario_cfg(name, N)
    terrain, heights = _terrain_of(state0, cfg)
    C = O.make_consts(cfg, cfg.sim.dt, robot_constants(cfg), terrain=terrain)
    S = O.load_state(C, O.new_state(C, N), state0)
    tight = dict(rtol=1e-6, atol=1e-7)      # same algorithm, same library; only libm vector width may differ
    for t, (inp, out) in enumerate(zip(inputs, outputs)):
        sim = SimpleNamespace(**{k: inp[k].clone() for k in SIM_KEYS})
        obs, priv, rew, reset, _ = O.step(C, S, sim, inp["actions"], pools_of(inp), terrain=terrain, height_samples=heights)
        assert torch.equal(reset, out["reset"]), f"{name} step {t}: reset flags"
        assert torch.equal(S.time_out_buf, out["time_out"])
        assert torch.equal(S.contact_filt, out["contact_filt"])
        torch.testing.assert_close(rew, out["rew"], **tight)
        torch.testing.assert_close(S.torques, out["torques"], **tight)
        torch.testing.assert_close(obs[:, -47:], out["obs_new"], **tight)
        torch.testing.assert_close(priv[:, -out["priv_new"].shape[1]:], out["priv_new"], **tight)
        torch.testing.assert_close(sim.root_states, out["root_after"], **tight)
        # disturbance windows (t1:193-247) and the command curriculum (lr:1160-1169)
        for k in ("applied_force", "applied_torque", "ext_forces", "ext_torques", "rand_push_force", "rand_push_torque"):
            torch.testing.assert_close(getattr(S, k), out[k], **tight)
        got = torch.tensor([S.command_ranges[k] for k in ("lin_vel_x", "lin_vel_y", "ang_vel_yaw")], dtype=torch.float64)
        assert torch.equal(got, out["command_ranges"]), f"{name} step {t}: command ranges"
    torch.testing.assert_close(obs, final["obs"], **tight)
    torch.testing.assert_close(priv, final["priv"], **tight)


def test_gae_oracle_matches_reference_fixture():
    z = np.load(os.path.join(GOLDEN, "gae_T24_N64.npz"))
    t = lambda k: torch.from_numpy(z[k])
    ret, adv = O.gae_returns(t("rewards"), t("values"), t("dones"), t("last_values"), float(z["gamma"]), float(z["lam"]))
    torch.testing.assert_close(ret, t("returns"), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(adv, t("advantages"), rtol=1e-6, atol=1e-7)


def test_reward_order_and_count():
    from ti5_isaacgym_b200.envs import DHT1StandCfg
    from ti5_isaacgym_b200.envs.t1.t1_robot import robot_constants
    cfg = DHT1StandCfg()
    C = O.make_consts(cfg, cfg.sim.dt, robot_constants(cfg))
    assert C.reward_names == sorted(C.reward_names) and len(C.reward_names) == 24


def test_stance_masks_do_not_depend_on_libm():
    """The gait phase takes 80 x 2 discrete values; none puts sin(2 pi phase) within 1e-4 of the
    0 / +-0.1 thresholds, so CPU and CUDA sinf (a few ulp apart) always give the same masks."""
    L = torch.arange(0, 2401)
    for start in (0.0, 0.5):
        phase = (L * 0.01 / 0.8) % 1.0 + start
        s = torch.sin(2 * torch.pi * phase.float()).double()
        near0 = s.abs() < 1e-4
        # the only near-zero values are the exact multiples of half a cycle, where |sin| < 0.1 makes both feet stance
        assert ((s.abs() - 0.1).abs() > 1e-4).all()
        assert (s[near0].abs() < 0.05).all()          # a sign flip there cannot change a mask


@pytest.mark.reference
def test_oracle_pinned_against_live_reference():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "oracle"))
    from oracle.pin_against_reference import SCENARIOS, pin_gae, run_scenario
    run_scenario("plane_events", dict(SCENARIOS["plane_events"], steps=8), None, verbose=False)
    run_scenario("trimesh_heights_push", dict(SCENARIOS["trimesh_heights_push"], steps=6), None, verbose=False)
    pin_gae(None)
