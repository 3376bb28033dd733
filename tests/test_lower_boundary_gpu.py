"""Lower boundary: the gym tensor-API calls one `step()` issues, compared call by call with what the
UNMODIFIED reference issued against the fake gym under oracle/shim for the same inputs (recorded in the golden
fixtures by oracle/pin_against_reference.py): names in order, the id lists and counts of the indexed setters, the
(N, 13, 3) force / torque tensors of apply_rigid_body_force_tensors, the pushed root tensor, and the per-env joint
property structs of lr:915-939.  Then the sync-free form of the same hand-over (device counts, one property tensor)."""
import numpy as np
import pytest
import torch

from helpers import close, exact, gym_calls_of, load_golden, pools_of, scenario_cfg, set_sim

pytestmark = pytest.mark.gpu


def _env(name, N, gym_cls):
    from ti5_isaacgym_b200.envs import T1DHStandEnv
    from ti5_isaacgym_b200.sim.synthetic import SimParams
    cfg = scenario_cfg(name, N)
    return T1DHStandEnv(cfg, SimParams(dt=cfg.sim.dt), 1, "cuda:0", True, gym=gym_cls(N, "cuda:0"), rng_mode="pools",
                        div_mode="ieee", use_cuda_graph=False, materialize_obs=False)


@pytest.mark.parametrize("name", ["plane_windows", "trimesh_windows", "plane_events"])
def test_gym_call_log_equals_the_references(name):
    from ti5_isaacgym_b200.sim.synthetic import RecordingGym
    state0, inputs, outputs, _ = load_golden(name)
    N = state0["commands"].shape[0]
    env = _env(name, N, RecordingGym)
    env.load_state(state0)
    n_force = n_push = n_props = 0
    for t, (inp, out) in enumerate(zip(inputs, outputs)):
        set_sim(env, inp)
        env.set_rng_pools(pools_of(inp))
        env.gym.calls.clear()
        env.step(inp["actions"].cuda())
        calls = list(env.gym.calls)
        tag = f"{name} step {t}: "
        assert [c[0] for c in calls] == gym_calls_of(out), tag + "gym call sequence differs from the reference's"
        ids = out["reset"].nonzero().flatten().to(torch.int32)
        for nm, payload in calls:
            if nm in ("set_dof_state_tensor_indexed", "set_actor_root_state_tensor_indexed"):
                state, got_ids, n = payload
                assert n == len(ids), tag + nm + " count"
                exact(got_ids, ids, tag + nm + " ids (ascending int32)")
                close(state, out["dof_after" if "dof" in nm else "root_after"], tag + nm + " state tensor")
            elif nm == "apply_rigid_body_force_tensors":
                f, tq, space = payload
                assert f.shape == (N, 13, 3) and tq.shape == (N, 13, 3) and space == 0
                assert float(f[:, 1:].abs().sum()) == 0 and float(tq[:, 1:].abs().sum()) == 0
                close(f[:, 0], out["applied_force"], tag + "force on the base")
                close(tq[:, 0], out["applied_torque"], tag + "torque on the base")
                n_force += int(bool(out["applied_force"].abs().sum() > 0))
            elif nm == "set_actor_root_state_tensor":
                # issued before the resets of the step, like t1:230: velocities are the drawn pushes
                close(payload[:, 7:9], out["rand_push_force"][:, :2], tag + "pushed linear velocity")
                close(payload[:, 10:13], out["rand_push_torque"], tag + "pushed angular velocity")
                n_push += 1
        props = [c[1] for c in calls if c[0] == "set_actor_dof_properties"]
        exact(torch.tensor([e for e, _ in props], dtype=torch.int64), out["props_env"], tag + "property structs: env order")
        if props:
            got = torch.stack([torch.from_numpy(d["armature"].copy()) for _, d in props])
            close(got, out["props_armature"], tag + "armatures handed to the simulator")
            n_props += len(props)
    assert n_props > 0
    if "windows" in name:
        assert n_force >= 5 and n_push >= 20


def test_sync_free_hand_over_with_device_counts():
    """A binding with `device_counts` receives (id buffer, device count) and one dense property tensor: same ids,
    same values, no host read-back inside step()."""
    from ti5_isaacgym_b200.sim.synthetic import SyntheticGym
    name = "plane_windows"
    state0, inputs, outputs, _ = load_golden(name)
    N = state0["commands"].shape[0]
    env = _env(name, N, SyntheticGym)
    env.load_state(state0)
    seen = 0
    for t, (inp, out) in enumerate(zip(inputs, outputs)):
        set_sim(env, inp)
        env.set_rng_pools(pools_of(inp))
        env.step(inp["actions"].cuda())
        (k1, ids1, n1), (k2, ids2, n2) = list(env.gym.indexed_calls)[-2:]
        assert (k1, k2) == ("dof", "root") and torch.is_tensor(n1) and n1.is_cuda and n1.dim() == 0
        ids = out["reset"].nonzero().flatten().to(torch.int32)
        assert int(n1) == len(ids) == int(n2)
        exact(ids1[:len(ids)], ids, f"step {t}: ids")
        p_ids, props, n = env.gym.dof_props
        assert props.shape == (N, 12, 3) and int(n) == len(ids)
        close(props[:len(ids), :, 2], out["props_armature"], f"step {t}: armatures, dense")
        assert bool((props[:len(ids), :, :2] == 1).all())
        seen += len(ids)
    assert seen > 0


def test_joint_friction_and_damping_multipliers_reach_the_simulator():
    """lr:755-773, 915-931 (`randomize_joint_friction` / `randomize_joint_damping`, off in t1_cfg): the multipliers of the
    re-spawned envs, drawn like the reference draws them (fixture recorded from the unmodified reference), in columns 0 / 1
    of the dense property tensor, and multiplied into the per-env property structs of an Isaac-Gym-like binding."""
    from ti5_isaacgym_b200.sim.synthetic import RecordingGym, SyntheticGym
    name = "plane_joint_props"
    state0, inputs, outputs, _ = load_golden(name)
    N = state0["commands"].shape[0]
    dense, rec = _env(name, N, SyntheticGym), _env(name, N, RecordingGym)
    for env in (dense, rec):
        env.load_state(state0)
    for e in range(N):                       # non-zero joint friction / damping to multiply into
        rec.gym._dof_props[e]["friction"][:] = 0.05 + 0.01 * np.arange(12)
        rec.gym._dof_props[e]["damping"][:] = 0.5 + 0.1 * np.arange(12)
    before = [rec.gym._dof_props[e].copy() for e in range(N)]
    n_props = 0
    for t, (inp, out) in enumerate(zip(inputs, outputs)):
        for env in (dense, rec):
            set_sim(env, inp)
            env.set_rng_pools(pools_of(inp))
        rec.gym.calls.clear()
        dense.step(inp["actions"].cuda()); rec.step(inp["actions"].cuda())
        ids = out["reset"].nonzero().flatten()
        tag = f"{name} step {t}: "
        close(dense.joint_friction_coeffs[ids], out["props_friction_coeff"], tag + "friction multipliers of the re-spawned envs")
        close(dense.joint_damping_coeffs[ids], out["props_damping_coeff"], tag + "damping multipliers")
        if len(ids):
            _, props, n = dense.gym.dof_props
            assert int(n) == len(ids)
            close(props[:len(ids), :, 0], out["props_friction_coeff"].expand(-1, 12), tag + "dense tensor, column 0")
            close(props[:len(ids), :, 1], out["props_damping_coeff"].expand(-1, 12), tag + "dense tensor, column 1")
            close(props[:len(ids), :, 2], out["props_armature"], tag + "dense tensor, column 2")
        structs = [c[1] for c in rec.gym.calls if c[0] == "set_actor_dof_properties"]
        exact(torch.tensor([e for e, _ in structs], dtype=torch.int64), out["props_env"], tag + "property structs: env order")
        for r, (e, d) in enumerate(structs):    # the struct the simulator held, times the multiplier (lr:923-930)
            close(torch.from_numpy(d["friction"].copy()), torch.from_numpy(before[e]["friction"]) * out["props_friction_coeff"][r],
                  tag + f"env {e}: friction")
            close(torch.from_numpy(d["damping"].copy()), torch.from_numpy(before[e]["damping"]) * out["props_damping_coeff"][r],
                  tag + f"env {e}: damping")
            before[e] = d.copy()
            n_props += 1
    assert n_props > 0


def test_refresh_actor_dof_props_for_an_explicit_id_list():
    """lr:915-939 called directly with env ids: ti5_gather_dof_props."""
    from ti5_isaacgym_b200.sim.synthetic import RecordingGym
    env = _env("plane_events", 64, RecordingGym)
    env.joint_armatures.copy_(torch.rand(64, 12, device="cuda"))
    ids = torch.tensor([3, 9, 10, 63], device="cuda")
    env.gym.calls.clear()
    env._refresh_actor_dof_props(ids)
    got = [(e, torch.from_numpy(d["armature"].copy())) for nm, (e, d) in env.gym.calls if nm == "set_actor_dof_properties"]
    assert [e for e, _ in got] == ids.tolist()
    exact(torch.stack([a for _, a in got]), env.joint_armatures[ids].cpu(), "armatures")
