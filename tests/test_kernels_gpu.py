"""Stand-alone kernels through the C ABI: GAE, reset-index compaction, height sampling; and the
properties of the production configuration (Philox RNG, CUDA graph, ring-buffer observation views)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, close, exact, make_env, scenario_cfg

from oracle import t1_oracle as O

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------- GAE (rs:97-119)
def _gae_case(T, N, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(T, N, 1, generator=g), torch.randn(T, N, 1, generator=g),
            (torch.rand(T, N, 1, generator=g) < 0.02).byte(), torch.randn(N, 1, generator=g))


@pytest.mark.parametrize("T,N", [(24, 64), (24, 8192), (5, 7), (1, 1), (24, 65536), (100, 129)])
def test_gae_matches_oracle(T, N):
    from ti5_isaacgym_b200.algo.rollout_storage import gae_returns_
    rew, val, done, last = _gae_case(T, N, 99)
    o_ret, o_adv = O.gae_returns(rew, val, done, last, 0.994, 0.9)
    ret, adv = torch.empty(T, N, 1, device="cuda"), torch.empty(T, N, 1, device="cuda")
    gae_returns_(rew.cuda(), val.cuda(), done.cuda(), last.cuda(), ret, adv, 0.994, 0.9)
    close(ret, o_ret, "returns", atol=1e-5)
    if T * N > 1:
        close(adv, o_adv, "advantages", atol=1e-5)


def test_gae_matches_reference_fixture_and_storage_class():
    from ti5_isaacgym_b200.algo.rollout_storage import RolloutStorage
    z = np.load(os.path.join(GOLDEN, "gae_T24_N64.npz"))
    st = RolloutStorage(64, 24, [4], [4], [2], device="cuda")
    st.rewards.copy_(torch.from_numpy(z["rewards"]))
    st.values.copy_(torch.from_numpy(z["values"]))
    st.dones.copy_(torch.from_numpy(z["dones"]))
    st.compute_returns(torch.from_numpy(z["last_values"]).cuda(), float(z["gamma"]), float(z["lam"]))
    close(st.returns, torch.from_numpy(z["returns"]), "returns vs reference", atol=1e-5)
    close(st.advantages, torch.from_numpy(z["advantages"]), "advantages vs reference", atol=1e-5)
    # size-independent property: normalised advantages have zero mean, unit (unbiased) std
    assert abs(float(st.advantages.mean())) < 1e-5 and abs(float(st.advantages.std()) - 1) < 1e-4


# ----------------------------------------------------------------------------- compaction (lr:490)
@pytest.mark.parametrize("n,rate", [(0, 0.5), (1, 1.0), (31, 0.3), (1024, 0.0), (1025, 1.0), (8192, 0.01), (65536, 0.05),
                                    (100000, 0.5)])
def test_compaction_is_ascending_nonzero(n, rate):
    from ti5_isaacgym_b200 import _lib
    lib = _lib.load_library()
    g = torch.Generator().manual_seed(n + 1)
    mask = (torch.rand(max(n, 1), generator=g) < rate)[:n].cuda()
    ids = torch.full((max(n, 1),), -1, dtype=torch.int32, device="cuda")
    count = torch.zeros(1, dtype=torch.int32, device="cuda")
    scratch = torch.zeros((n + 1023) // 1024 + 2, dtype=torch.int32, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    for _ in range(2):       # twice: the scratch ticket must be left clean
        _lib.check(lib.ti5_compact_resets(p(mask), n, p(ids), p(count), p(scratch), None))
    want = mask.nonzero().flatten().to(torch.int32)
    assert int(count) == len(want)
    exact(ids[:len(want)], want, "compacted ids")


# ----------------------------------------------------------------------------- production configuration
def _production_env(N, name="plane_default", **kw):
    from ti5_isaacgym_b200.envs import T1DHStandEnv
    from ti5_isaacgym_b200.sim.synthetic import SimParams, fill_synthetic_state
    cfg = scenario_cfg(name, N)
    kw.setdefault("materialize_obs", False)
    env = T1DHStandEnv(cfg, SimParams(dt=cfg.sim.dt), 1, "cuda:0", True, seed=11, **kw)
    gen = torch.Generator(device="cuda").manual_seed(5)
    fill_synthetic_state(env.gym.tensors, env.env_origins, gen, base_contact_rate=0.03)
    return env, gen


def test_graph_replay_equals_direct_launches_and_history_layout():
    from ti5_isaacgym_b200.sim.synthetic import synthetic_actions
    N = 2048
    torch.manual_seed(0)
    a, gen_a = _production_env(N, use_cuda_graph=True)
    torch.manual_seed(0)
    b, gen_b = _production_env(N, use_cuda_graph=False, materialize_obs=True)
    a.reset(), b.reset()
    prev = None
    for t in range(70):
        act = synthetic_actions(N, gen_a, "cuda")
        oa, pa, ra, da, xa = a.step(act)
        ob, pb, rb, db, xb = b.step(act)
        exact(oa, ob, f"step {t}: graph vs direct obs (ring view vs materialised copy)")
        exact(pa, pb, f"step {t}: privileged obs")
        exact(ra, rb, f"step {t}: rewards")
        exact(da, db, f"step {t}: resets")
        assert oa.shape == (N, 66 * 47) and oa.stride() == (2 * 66 * 47, 1) and ob.is_contiguous()
        if prev is not None:
            keep = ~da                                      # history of a re-spawned env is cleared (t1:556-559)
            exact(oa[keep][:, :-47], prev[keep][:, 47:], f"step {t}: history shifts by one frame")
            assert (oa[da][:, :-47] == 0).all()
        prev = oa.clone()
        assert oa.view(-1, 66, 47).shape == (N, 66, 47) and oa[..., -235:].shape == (N, 235)    # actor_critic_dh.py:154-159
    assert float(oa.abs().max()) <= 100.0 and torch.isfinite(oa).all()
    g = a.sync_from_device()
    assert g.step_index == b.sync_from_device().step_index


@pytest.mark.parametrize("N,graph,fused", [(8192, True, True), (1000, False, True), (16384, True, True), (8192, True, False),
                                           (1000, False, False), (65536, True, True), (65536, True, False)])
def test_chained_launches_equal_plain_launches(N, graph, fused):
    """Programmatic dependent launches (TI5_*_CHAINED) only move work in front of the grid wait: every output and
    every piece of state must be bit-identical to ordinary stream-ordered launches, resets included."""
    from ti5_isaacgym_b200.sim.synthetic import synthetic_actions
    torch.manual_seed(0)                                    # friction / mass draws at construction
    a, gen = _production_env(N, use_cuda_graph=graph, chain_launches=True, fused_step=fused)
    torch.manual_seed(0)
    b, _ = _production_env(N, use_cuda_graph=graph, chain_launches=False, fused_step=fused)
    a.reset(), b.reset()
    ep = torch.randint(1, 2400, (N,), generator=gen, device="cuda")
    ep[:8] = 2399                                           # time-outs in the first steps
    a.episode_length_buf, b.episode_length_buf = ep.clone(), ep.clone()
    n_reset = 0
    for t in range(12 if N >= 65536 else 40):
        act = synthetic_actions(N, gen, "cuda")
        oa, pa, ra, da, _ = a.step(act)
        ob, pb, rb, db, _ = b.step(act)
        exact(da, db, f"step {t}: resets")
        exact(ra, rb, f"step {t}: rewards")
        exact(oa, ob, f"step {t}: obs")
        exact(pa, pb, f"step {t}: privileged obs")
        n_reset += int(da.sum())
    assert n_reset > 0
    for name in ("torques", "commands", "episode_length_buf", "dof_state", "root_states", "last_actions", "feet_air_time",
                 "base_lin_vel", "base_euler_xyz", "lag_buffer", "dof_lag_buffer", "imu_lag_buffer"):
        exact(getattr(a, name), getattr(b, name), f"state after 40 steps: {name}")
    exact(a._episode_sums, b._episode_sums, "episode sums")


@pytest.mark.parametrize("N,graph,name,chain", [
    (8192, True, "plane_default", True), (1000, False, "plane_default", True), (4113, False, "plane_default", False),
    (65536, True, "plane_default", True),                    # several waves of CTAs
    (8192, True, "trimesh_heights_push", True), (517, False, "trimesh_windows", True)])
def test_fused_step_equals_the_twelve_launch_sequence(N, graph, name, chain):
    """ti5_fused_step (clip + DEC substeps + post-physics in one launch) against ti5_first_substep, 9 x ti5_substep,
    ti5_post_physics: same Philox counters, same op order -> every output and every piece of state bit-identical."""
    from ti5_isaacgym_b200.sim.synthetic import synthetic_actions
    torch.manual_seed(0)
    a, gen = _production_env(N, name, use_cuda_graph=graph, fused_step=True, chain_launches=chain)
    torch.manual_seed(0)
    b, _ = _production_env(N, name, use_cuda_graph=graph, fused_step=False, chain_launches=chain)
    assert a.launches_per_step + 10 == b.launches_per_step
    a.reset(), b.reset()
    ep = torch.randint(1, 2400, (N,), generator=gen, device="cuda")
    ep[:8] = 2399                                           # time-outs in the first steps
    a.episode_length_buf, b.episode_length_buf = ep.clone(), ep.clone()
    if "windows" in name:
        a.set_common_step_counter(287997), b.set_common_step_counter(287997)
    n_reset = 0
    for t in range(24 if N >= 65536 else 40):
        act = synthetic_actions(N, gen, "cuda")
        oa, pa, ra, da, _ = a.step(act)
        ob, pb, rb, db, _ = b.step(act)
        exact(da, db, f"step {t}: resets")
        exact(ra, rb, f"step {t}: rewards")
        exact(oa, ob, f"step {t}: obs")
        exact(pa, pb, f"step {t}: privileged obs")
        exact(a.torques, b.torques, f"step {t}: torques")
        exact(a.torque_multi, b.torque_multi, f"step {t}: torque multipliers")
        n_reset += int(da.sum())
    assert n_reset > 0
    for nm in ("actions", "commands", "episode_length_buf", "dof_state", "root_states", "last_actions", "feet_air_time",
               "base_lin_vel", "base_euler_xyz", "lag_buffer", "dof_lag_buffer", "imu_lag_buffer", "applied_force",
               "rand_push_torque", "ext_forces", "reset_ids"):
        exact(getattr(a, nm), getattr(b, nm), f"state after the run: {nm}")
    exact(a._episode_sums, b._episode_sums, "episode sums")
    exact(a._reward_terms, b._reward_terms, "reward terms")
    assert int(a.sync_from_device().step_index) == int(b.sync_from_device().step_index)


@pytest.mark.parametrize("name", ["plane_lag_perstep", "plane_pos_vel_lag"])
def test_lag_options_in_production_mode(name):
    """The lag options t1_cfg leaves off with in-kernel Philox draws (parity with the reference is established in pools
    mode): every lag index stays in its range, a per-step re-draw never looks more than one step further back than the
    last one (lr:1041-1042, t1:411-412), the indices do move, and the CUDA-graph step equals direct launches."""
    from ti5_isaacgym_b200.sim.synthetic import synthetic_actions
    N = 1024
    torch.manual_seed(0)
    a, gen = _production_env(N, name, use_cuda_graph=True)
    torch.manual_seed(0)
    b, _ = _production_env(N, name, use_cuda_graph=False)
    a.reset(), b.reset()
    dr = a.cfg.domain_rand
    cols = {"lag_timestep": dr.lag_timesteps_range, "imu_lag_timestep": dr.imu_lag_timesteps_range}
    if name == "plane_pos_vel_lag":
        cols.update(dof_pos_lag_timestep=dr.dof_pos_lag_timesteps_range, dof_vel_lag_timestep=dr.dof_vel_lag_timesteps_range)
        perstep = ["dof_pos_lag_timestep"]
    else:
        cols.update(dof_lag_timestep=dr.dof_lag_timesteps_range)
        perstep = ["dof_lag_timestep", "imu_lag_timestep"]       # (the action lag moves ten times per step)
    prev = {k: getattr(a, k).clone() for k in cols}
    moved = {k: 0 for k in perstep}
    for t in range(40):
        act = synthetic_actions(N, gen, "cuda")
        oa, pa, ra, da, _ = a.step(act)
        ob, pb, rb, db, _ = b.step(act.clone())
        exact(oa, ob, f"{name} step {t}: obs (graph vs direct)"); exact(ra, rb, f"step {t}: rewards"); exact(da, db, f"step {t}: resets")
        for k, (lo, hi) in cols.items():
            v = getattr(a, k)
            exact(v, getattr(b, k), f"step {t}: {k}")
            assert int(v.min()) >= lo and int(v.max()) <= hi, f"step {t}: {k} outside [{lo}, {hi}]"
        for k in perstep:
            v = getattr(a, k)
            ok = (v <= prev[k] + 1) | da            # a re-spawned env starts over from the range maximum
            assert bool(ok.all()), f"step {t}: {k} jumped back by more than one step"
            moved[k] += int((v != prev[k]).sum())
        prev = {k: getattr(a, k).clone() for k in cols}
    assert all(m > N for m in moved.values()), moved


def test_use_ref_actions_offsets_the_policy_output():
    """t1:360-366 `env.use_ref_actions`: the step runs on actions + ref_action (added in place, as the reference does)."""
    from ti5_isaacgym_b200.sim.synthetic import synthetic_actions
    N = 512
    envs = []
    for flag in (True, False):
        torch.manual_seed(0)
        env, gen = _production_env(N, use_cuda_graph=False)
        env.cfg.env.use_ref_actions = flag
        env.reset()
        envs.append(env)
    a, b = envs
    for t in range(6):
        act = synthetic_actions(N, gen, "cuda")
        want = act + b.ref_action
        given = act.clone()
        oa, _, ra, da, _ = a.step(given)
        ob, _, rb, db, _ = b.step(want.clone())
        exact(given, want, f"step {t}: the caller's tensor carries the offset")
        exact(oa, ob, f"step {t}: obs"); exact(ra, rb, f"step {t}: rewards"); exact(da, db, f"step {t}: resets")
        exact(a.ref_action, b.ref_action, f"step {t}: reference action")


def test_graph_cache_follows_the_action_buffer():
    """The captured step reads the caller's action tensor in place (one graph per buffer address, at most 8); beyond
    that, and for host / strided inputs, it goes through the static copy.  Same results as direct launches."""
    from ti5_isaacgym_b200.sim.synthetic import synthetic_actions
    N = 1024
    torch.manual_seed(0)
    a, gen = _production_env(N, use_cuda_graph=True)
    torch.manual_seed(0)
    b, _ = _production_env(N, use_cuda_graph=False)
    a.reset(), b.reset()
    bufs = [synthetic_actions(N, gen, "cuda") for _ in range(12)]          # 12 live buffers: 8 graphs + the fallback
    wide = torch.zeros(N, 24, device="cuda")
    for t in range(30):
        act = bufs[t % 12]
        if t % 5 == 4:
            wide[:, ::2] = act
            given = wide[:, ::2]                                           # strided view: copied
        elif t % 7 == 6:
            given = act.cpu()                                              # host tensor: copied
        else:
            given = act
        oa, pa, ra, da, _ = a.step(given)
        ob, pb, rb, db, _ = b.step(act)
        exact(oa, ob, f"step {t}: obs"); exact(ra, rb, f"step {t}: rewards"); exact(da, db, f"step {t}: resets")
    assert len(a._graphs) == a._max_graphs == 8 and a._graph is not None


def test_tiling_does_not_change_results(monkeypatch):
    """env_block 32 (writer warps, reset draws parked in shared memory before the grid wait) and env_block 64 (neither)
    are two schedules of the same arithmetic: bit-identical step outputs, resets included."""
    from ti5_isaacgym_b200.sim.synthetic import synthetic_actions
    N = 4096 + 17
    envs = []
    for tb in ("32", "64"):
        monkeypatch.setenv("TI5_ENV_BLOCK", tb)
        torch.manual_seed(0)
        env, gen = _production_env(N, use_cuda_graph=False)
        assert env._params.env_block == int(tb)
        env.reset()
        envs.append(env)
    a, b = envs
    ep = torch.randint(1, 2400, (N,), generator=gen, device="cuda")
    ep[-5:] = 2399
    a.episode_length_buf, b.episode_length_buf = ep.clone(), ep.clone()
    n_reset = 0
    for t in range(30):
        act = synthetic_actions(N, gen, "cuda")
        oa, pa, ra, da, _ = a.step(act)
        ob, pb, rb, db, _ = b.step(act)
        exact(da, db, f"step {t}: resets")
        exact(ra, rb, f"step {t}: rewards")
        exact(oa, ob, f"step {t}: obs")
        exact(pa, pb, f"step {t}: privileged obs")
        n_reset += int(da.sum())
    assert n_reset > 20
    for name in ("dof_state", "root_states", "commands", "gait_time", "lag_timestep", "motor_offsets", "randomized_p_gains"):
        exact(getattr(a, name), getattr(b, name), f"state after 30 steps: {name}")


def test_full_size_properties_8192():
    """BASELINE size: properties that need no oracle."""
    from ti5_isaacgym_b200.algo.vec_env import check_vec_env
    from ti5_isaacgym_b200.sim.synthetic import synthetic_actions
    N = 8192
    env, gen = _production_env(N)
    env.reset()
    assert check_vec_env(env) == [], check_vec_env(env)
    env.episode_length_buf = torch.randint(1, 2400, (N,), generator=gen, device="cuda")
    for t in range(30):
        act = synthetic_actions(N, gen, "cuda")
        ep_before = env.episode_length_buf.clone()
        obs, priv, rew, done, extras = env.step(act)
        g = env.sync_from_device()
        ids = env.reset_ids[:g.n_reset].long()
        exact(ids, done.nonzero().flatten(), "reset ids == nonzero(reset_buf), ascending")
        assert (rew >= 0).all() and torch.isfinite(rew).all()
        total = sum(env.reward_terms[n] for n in env.reward_names)
        close(rew, total.clamp(min=0), "rew = clip(sum of scaled terms)", atol=1e-5)
        assert (env.episode_length_buf[done] == 0).all()
        exact(env.episode_length_buf[~done], ep_before[~done] + 1, "episode counters")
        assert (env.torques.abs() <= env.torque_limits + 1e-4).all()
        tm = env.torque_multi
        assert 0.8 <= float(tm.min()) and float(tm.max()) <= 1.2 and abs(float(tm.mean()) - 1.0) < 5e-3
        close(obs[:, -47 + 29:-47 + 41], env.actions, "newest frame carries the clipped actions")
        assert set(extras["episode"].keys()) == {"rew_" + n for n in env.reward_names} | {"max_command_x"}
    assert priv.shape == (N, 219)


def test_episode_length_setter_feeds_the_kernels():
    env, gen = _production_env(256)
    env.reset()
    env.episode_length_buf = torch.full((256,), 2400, device="cuda")       # dh_on_policy_runner.py:101 rebinding
    _, _, _, done, extras = env.step(torch.zeros(256, 12, device="cuda"))
    assert done.all() and extras["time_outs"].all()


def test_heights_match_oracle_at_scale():
    from ti5_isaacgym_b200.envs.t1.t1_robot import robot_constants
    from ti5_isaacgym_b200.sim.synthetic import fill_synthetic_state
    N = 4096
    cfg = scenario_cfg("trimesh_heights_push", N)
    env = make_env(cfg, div_mode="reciprocal")
    gen = torch.Generator(device="cuda").manual_seed(9)
    fill_synthetic_state(env.gym.tensors, env.env_origins, gen)
    got = env._get_heights().clone()
    C = O.make_consts(cfg, cfg.sim.dt, robot_constants(cfg), device="cuda:0")
    S = O.new_state(C, N)
    S.base_quat = env.root_states[:, 3:7].clone()
    want = O.sample_heights(C, S, env.gym.tensors, env.height_samples)
    # The cell index is a truncation of an fp32 coordinate.  torch's CUDA `norm` reduction inside quat_apply_yaw may
    # contract x*x + acc into an FMA where this library (-fmad=false) rounds twice, so the coordinate can differ in its
    # last bit; that changes the cell ONLY when the coordinate sits within one ulp of a cell edge.  Everywhere else the
    # sample must be the same cell, i.e. the same height, bit for bit.
    coords = O.height_coords(C, S, env.gym.tensors)[..., :2]                       # (N, npts, 2) before .long()
    ulp = torch.nextafter(coords.abs(), torch.full_like(coords, float("inf"))) - coords.abs()
    on_edge = ((coords - coords.round()).abs() <= ulp).any(-1)
    differ = got != want
    assert not (differ & ~on_edge).any(), f"{int((differ & ~on_edge).sum())} samples landed in a different cell away from any cell edge"
    assert float(differ.float().mean()) < 2e-4
    exact(got[~differ], want[~differ], "heights")


def test_host_io_graph_equals_the_device_path():
    """enable_host_io(): actions read from pinned host memory by a copy node inside the step graph, results landed
    in pinned host memory — same step as `step(actions)` on device tensors."""
    from ti5_isaacgym_b200.sim.synthetic import synthetic_actions
    N = 1000
    torch.manual_seed(0)
    a, gen_a = _production_env(N, use_cuda_graph=True)
    torch.manual_seed(0)
    b, _ = _production_env(N, use_cuda_graph=True)
    a.reset(), b.reset()
    h_act, h_out = b.enable_host_io()
    assert h_act.is_pinned() and h_out.is_pinned() and h_act.shape == (N, 12)
    for t in range(12):
        act = synthetic_actions(N, gen_a, "cuda")
        oa, pa, ra, da, xa = a.step(act)
        h_act.copy_(act.cpu())
        ob, pb, rb, db, xb = b.step_host()
        exact(oa, ob, f"step {t}: obs"); exact(pa, pb, f"step {t}: privileged obs")
        # the host copies are complete when step_host returns: no further synchronisation here
        assert torch.equal(b.host_rew, ra.cpu()) and torch.equal(b.host_reset, da.cpu())
        assert torch.equal(b.host_time_outs, xa["time_outs"].cpu())
    assert int(b.sync_from_device().step_index) == int(a.sync_from_device().step_index)
