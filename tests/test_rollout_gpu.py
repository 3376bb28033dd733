"""Frame-log rollout storage (ti5_store_transition / ti5_gather_minibatch) against the reference's storage:
the committed fixture of the reference's own outputs, the oracle on larger random rollouts, and the real env
in the loop (the env's frame logs vs. copies of the windows it returned)."""
from types import SimpleNamespace

import pytest
import torch

from helpers import exact, make_env, scenario_cfg
from oracle import rollout_oracle as RO
from test_rollout_cpu import COLS, load_rollout_golden, windows_from_frames

pytestmark = pytest.mark.gpu


class FrameSource:
    """Stands in for the env on the storage's side of the boundary: the frame logs filled exactly as
    ti5_reset_observe fills them (row (step-1) % L; valid = frames since the last clear, capped at H)."""

    def __init__(self, pre_obs, pre_priv, num_actions=12, offset=0):
        self.device = torch.device("cuda:0")
        self.H, self.CH = pre_obs.shape[0], pre_priv.shape[0]
        self.num_envs, self.K, self.P = pre_obs.shape[1], pre_obs.shape[2], pre_priv.shape[2]
        self.num_obs, self.num_privileged_obs, self.num_actions = self.H * self.K, self.CH * self.P, num_actions
        self._pre, self._step, self.L = (pre_obs, pre_priv), offset, 0

    def enable_frame_log(self, T):
        N, dev = self.num_envs, self.device
        self.L = L = T + max(self.H, self.CH) + 2
        self.frame_log, self.priv_log = torch.zeros(N, L, self.K, device=dev), torch.zeros(N, L, self.P, device=dev)
        self.valid_log = torch.zeros(L, N, dtype=torch.int16, device=dev)
        self.hist_valid = torch.full((N,), self.H, dtype=torch.int32, device=dev)
        for log, pre in zip((self.frame_log, self.priv_log), self._pre):
            n = pre.shape[0]
            rows = (self._step - n + torch.arange(n)) % L
            log[:, rows] = pre.permute(1, 0, 2).to(dev)
        self.valid_log[(self._step - 1) % L] = self.H

    def frame_logs(self):
        return SimpleNamespace(frame_stack=self.H, c_frame_stack=self.CH, num_single_obs=self.K, priv_frame=self.P,
                               log_len=self.L, frame_log=self.frame_log, priv_log=self.priv_log, valid_log=self.valid_log)

    @property
    def frame_log_row(self):
        return (self._step - 1) % self.L

    def env_step(self, frame, priv_frame, dones):
        self._step += 1
        row = (self._step - 1) % self.L
        self.frame_log[:, row], self.priv_log[:, row] = frame.to(self.device), priv_frame.to(self.device)
        hv = torch.where(dones.to(self.device), 0, self.hist_valid)
        self.hist_valid = torch.clamp(hv + 1, max=self.H).int()
        self.valid_log[row] = self.hist_valid.short()


def _fill(storage, src, R, gamma, fused, tag_rows=True):
    from ti5_isaacgym_b200.algo.rollout_storage import RolloutStorage
    T = R["rewards"].shape[0]
    for t in range(T):
        tr = RolloutStorage.Transition()
        obs = torch.empty(0, device="cuda")                          # content is never read: the logs hold the frames
        if tag_rows:
            obs.ti5_frame_row = src.frame_log_row
        tr.observations = obs
        tr.actions, tr.values, tr.actions_log_prob = R["actions"][t].cuda(), R["values"][t].cuda(), R["log_prob"][t].cuda()
        tr.action_mean, tr.action_sigma = R["mean"][t].cuda(), R["sigma"][t].cuda()
        src.env_step(R["obs_frames"][t], R["priv_frames"][t], R["dones"][t])
        if fused:
            storage.store_step(tr, R["rewards"][t].cuda(), R["dones"][t].cuda(), R["time_outs"][t].cuda(), gamma)
        else:
            tr.rewards = RO.bootstrap_rewards(R["rewards"][t], R["values"][t], R["time_outs"][t], gamma).cuda()
            tr.dones = R["dones"][t].cuda()
            storage.add_transitions(tr)


def test_golden_rollout_of_the_reference():
    from ti5_isaacgym_b200.algo.rollout_storage import FrameLogRolloutStorage
    R, gamma, out = load_rollout_golden()
    T, N = R["rewards"].shape
    src = FrameSource(R["pre_obs"], R["pre_priv"])
    st = FrameLogRolloutStorage(src, T)
    _fill(st, src, R, gamma, fused=True)
    exact(st.rewards, out["rewards"], "rewards + gamma * values * time_outs (dh_ppo.py:97-98)")
    exact(st.dones, out["dones"], "dones")
    exact(st.observations, out["observations"], "observations (T,N,H*K) rebuilt from the frame log")
    exact(st.privileged_observations, out["privileged_observations"], "privileged_observations")
    rb, lb = st.finished_episodes()
    assert rb == out["rewbuffer"].tolist() and lb == out["lenbuffer"].tolist()
    st.returns.copy_(R["returns"].cuda()); st.advantages.copy_(R["advantages"].cuda())
    batches = list(st.mini_batch_generator(4, 1, indices=R["indices"].cuda()))
    assert len(batches) == 4 and batches[0][9] == (None, None) and batches[0][10] is None
    for i, b in enumerate(batches):
        for c, v in zip(COLS, b[:9]):
            exact(v, out[f"batch{i}.{c}"], f"batch {i} {c}")
    with pytest.raises(AssertionError, match="Rollout buffer overflow"):
        _fill(st, src, {k: v[:1] if torch.is_tensor(v) and v.shape[0] == T else v for k, v in R.items()}, gamma, fused=False)


@pytest.mark.parametrize("T,N,H,CH,P,fused,offset,nmb", [
    (24, 300, 66, 3, 73, True, 0, 4),        # the BASELINE rollout shape, N not a multiple of the CTA tile
    (24, 257, 66, 3, 260, False, 1000, 3),   # measure_heights frames; the log wraps (offset) ; plain add_transitions
    (7, 1, 2, 1, 73, True, 5, 1),            # a single env, minimal stacks
    (9, 130, 15, 3, 73, True, 77, 5),
])
def test_random_rollouts_follow_the_oracle(T, N, H, CH, P, fused, offset, nmb):
    from ti5_isaacgym_b200.algo.rollout_storage import FrameLogRolloutStorage
    R = RO.synthetic_rollout(T, N, H, CH, 47, P, 12, seed=T * 1000 + N)
    gamma = 0.994
    mb = (T * N) // nmb
    idx = torch.randperm(nmb * mb, generator=torch.Generator().manual_seed(1))
    S, rb, lb, ref_batches = RO.run_rollout(R, gamma, nmb, 2, idx)
    src = FrameSource(R["pre_obs"], R["pre_priv"], offset=offset)
    st = FrameLogRolloutStorage(src, T)
    _fill(st, src, R, gamma, fused, tag_rows=fused)      # the un-tagged path uses the runner's call protocol
    exact(st.rewards, S.rewards, "rewards")
    if fused:
        r2, l2 = st.finished_episodes()
        assert r2 == rb and l2 == lb
        exact(st.cur_reward_sum, RO_cur(R)[0], "cur_reward_sum"); exact(st.cur_episode_length, RO_cur(R)[1], "cur_episode_length")
    got = list(st.mini_batch_generator(nmb, 2, indices=idx.cuda()))
    assert len(got) == len(ref_batches)
    for i, (g, w) in enumerate(zip(got, ref_batches)):
        for c, u, v in zip(COLS, g[:9], w):
            exact(u, v, f"batch {i} {c}")
    # GAE on the stored rewards / values / dones (rs:97-119) still runs on this storage
    st.compute_returns(torch.zeros(N, 1, device="cuda"), gamma, 0.9)
    assert torch.isfinite(st.advantages).all()


def RO_cur(R):
    N = R["rewards"].shape[1]
    cs, cl = torch.zeros(N), torch.zeros(N)
    for t in range(R["rewards"].shape[0]):
        RO.episode_bookkeeping(cs, cl, R["rewards"][t], R["dones"][t], [], [])
    return cs, cl


@pytest.mark.parametrize("use_graph,H", [(False, 66), (True, 66), (True, 5)])
def test_env_frame_logs_reproduce_the_windows_the_env_returned(use_graph, H):
    """The real env in the loop, with resets: the windows rebuilt from the logs ti5_reset_observe writes equal the
    windows the env handed to the policy (cloned at act time, which is what the reference stores, rs:62-63)."""
    from ti5_isaacgym_b200.algo.rollout_storage import FrameLogRolloutStorage, RolloutStorage
    from ti5_isaacgym_b200.envs import T1DHStandEnv
    from ti5_isaacgym_b200.sim.synthetic import SimParams, fill_synthetic_state, synthetic_actions
    N, T = 384, 24
    cfg = scenario_cfg("plane_default", N, frame_stack=H)
    env = T1DHStandEnv(cfg, SimParams(dt=cfg.sim.dt), 1, "cuda:0", True, rng_mode="philox", div_mode="reciprocal",
                       use_cuda_graph=use_graph, seed=11)
    gen = torch.Generator(device="cuda").manual_seed(5)
    obs, priv = env.reset()
    for _ in range(3):                                               # some history before the storage is attached
        fill_synthetic_state(env.gym.tensors, env.env_origins, gen, base_contact_rate=0.05)
        obs, priv, *_ = env.step(synthetic_actions(N, gen, "cuda"))
    fl = FrameLogRolloutStorage(env, T)
    plain = RolloutStorage(N, T, [env.num_obs], [env.num_privileged_obs], [12], None, "cuda")
    n_done = 0
    for rollout in range(3):                                         # the log rows wrap across rollouts
        fl.clear(); plain.clear()
        for t in range(T):
            tr, tp = RolloutStorage.Transition(), RolloutStorage.Transition()
            for x in (tr, tp):
                x.actions = synthetic_actions(N, gen, "cuda") if x is tr else tr.actions
                x.values, x.actions_log_prob = torch.full((N, 1), 0.25 * t, device="cuda"), torch.full((N,), -1.0 * t, device="cuda")
                x.action_mean, x.action_sigma = x.actions * 0.5, x.actions.abs() + 0.1
            tr.observations = obs                                    # the (tagged) view, read one step late
            tp.observations, tp.critic_observations = obs.clone(), priv.clone()
            fill_synthetic_state(env.gym.tensors, env.env_origins, gen, base_contact_rate=0.05)
            obs, priv, rew, dones, infos = env.step(tr.actions)
            n_done += int(dones.sum())
            fl.store_step(tr, rew, dones, infos["time_outs"], 0.994)
            tp.rewards = RO.bootstrap_rewards(rew, tp.values, infos["time_outs"], 0.994)
            tp.dones = dones
            plain.add_transitions(tp)
        exact(fl.rewards, plain.rewards, "rewards"); exact(fl.dones, plain.dones, "dones")
        exact(fl.observations, plain.observations, f"observations, rollout {rollout}")
        exact(fl.privileged_observations, plain.privileged_observations, f"privileged observations, rollout {rollout}")
        idx = torch.randperm(T * N, device="cuda", generator=gen)
        for g, w in zip(fl.mini_batch_generator(4, 1, indices=idx), _plain_batches(plain, 4, idx)):
            for c, u, v in zip(COLS, g[:9], w):
                exact(u, v, c)
    assert n_done > 50                                               # the run did exercise the cleared windows


def _plain_batches(st, nmb, idx):
    flat = lambda t: t.flatten(0, 1)
    mb = idx.numel() // nmb
    cols = [flat(t) for t in (st.observations, st.privileged_observations, st.actions, st.values, st.advantages, st.returns,
                              st.actions_log_prob, st.mu, st.sigma)]
    for i in range(nmb):
        yield tuple(c[idx[i * mb:(i + 1) * mb]] for c in cols)


def test_held_ring_view_is_not_a_stored_observation():
    """Why the storage reads the logs: the (N, H*K) tensor `step()` returns is a view into the env's history ring,
    and the runner keeps it across the NEXT step before storing it (dh_ppo.py:88 -> runner :136 -> rs:62)."""
    N = 256
    cfg = scenario_cfg("plane_default", N)
    env = make_env(cfg, rng_mode="philox", div_mode="reciprocal")
    from ti5_isaacgym_b200.sim.synthetic import fill_synthetic_state, synthetic_actions
    gen = torch.Generator(device="cuda").manual_seed(1)
    fill_synthetic_state(env.gym.tensors, env.env_origins, gen)
    obs, *_ = env.step(synthetic_actions(N, gen, "cuda"))
    kept = obs.clone()
    env.step(synthetic_actions(N, gen, "cuda"))
    assert not torch.equal(obs, kept)              # the view moved on: only a copy, or the frame log, preserves it
    # materialize_obs=True hands out fresh tensors like the reference (t1:477-481): they outlive later steps
    env2 = make_env(scenario_cfg("plane_default", N), rng_mode="philox", div_mode="reciprocal", materialize_obs=True)
    fill_synthetic_state(env2.gym.tensors, env2.env_origins, gen)
    o2, p2, *_ = env2.step(synthetic_actions(N, gen, "cuda"))
    assert o2.is_contiguous() and p2.is_contiguous()
    exact(o2, env2._history_views()[0], "materialised window")
    k2, kp2 = o2.clone(), p2.clone()
    for _ in range(3):
        env2.step(synthetic_actions(N, gen, "cuda"))
    exact(o2, k2, "held observation"); exact(p2, kp2, "held privileged observation")


def test_collector_runs_the_runner_loop_without_per_step_syncs():
    """runner :130-172 through RolloutCollector with a stand-in PPO object: what lands in the storage equals what
    the reference's process_env_step / add_transitions / bookkeeping produce from the same policy outputs."""
    from ti5_isaacgym_b200.algo.collector import RolloutCollector
    from ti5_isaacgym_b200.algo.rollout_storage import RolloutStorage
    from ti5_isaacgym_b200.envs import T1DHStandEnv
    from ti5_isaacgym_b200.sim.synthetic import SimParams, fill_synthetic_state
    N, T = 200, 12
    cfg = scenario_cfg("plane_default", N, frame_stack=8)
    env = T1DHStandEnv(cfg, SimParams(dt=cfg.sim.dt), 1, "cuda:0", True, rng_mode="philox", div_mode="reciprocal", seed=3)
    gen = torch.Generator(device="cuda").manual_seed(2)
    seen = []

    class Alg:                                            # the members of DHPPO the loop touches (dh_ppo.py:76-110)
        gamma, storage = 0.994, None
        transition = RolloutStorage.Transition()
        actor_critic = SimpleNamespace(reset=lambda dones: None)

        def act(self, obs, critic_obs):
            tr = self.transition
            tr.actions = torch.tanh(obs[:, -47:-35]) + 0.1 * torch.randn(N, 12, device="cuda", generator=gen)
            tr.values = critic_obs[:, :1] * 0.5
            tr.actions_log_prob = -tr.actions.square().sum(1)
            tr.action_mean, tr.action_sigma = tr.actions * 0.9, torch.full((N, 12), 0.3, device="cuda")
            tr.observations, tr.critic_observations = obs, critic_obs
            seen.append((obs.clone(), critic_obs.clone(), tr.actions, tr.values, tr.actions_log_prob))
            fill_synthetic_state(env.gym.tensors, env.env_origins, gen, base_contact_rate=0.08)   # next physics state
            return tr.actions

        def compute_returns(self, last_critic_obs):
            self.storage.compute_returns(last_critic_obs[:, :1] * 0.5, self.gamma, 0.9)

    # the stand-in answers to the signatures of the reference's DHPPO (recorded from the unmodified tree by
    # oracle/pin_signatures.py; the GPU box has no reference)
    import inspect
    import json
    import os
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_signatures.json")))
    for meth in ("act", "compute_returns"):
        assert list(inspect.signature(getattr(Alg, meth)).parameters) == ref[f"DHPPO.{meth}"], meth
    env.reset()
    alg = Alg()
    col = RolloutCollector(env, alg, T)
    assert list(inspect.signature(alg.process_env_step).parameters) == ref["DHPPO.process_env_step"][1:]
    want_rb, want_lb = [], []
    cs, cl = torch.zeros(N), torch.zeros(N)
    for rollout in range(2):
        seen.clear()
        st = col.collect()
        assert st.step == T and len(seen) == T
        exact(st.observations, torch.stack([s[0] for s in seen]), "stored observations = what the policy saw")
        exact(st.privileged_observations, torch.stack([s[1] for s in seen]), "stored critic observations")
        exact(st.actions, torch.stack([s[2] for s in seen]), "actions")
        exact(st.actions_log_prob.squeeze(-1), torch.stack([s[4] for s in seen]), "log prob")
        assert torch.isfinite(st.returns).all() and abs(float(st.advantages.mean())) < 1e-3
        # no env reaches the 2400-step time-out here, so the stored rewards are the env's (dh_ppo.py:97-98 adds 0)
        for t in range(T):
            RO.episode_bookkeeping(cs, cl, st.rewards[t, :, 0].cpu(), st.dones[t, :, 0].cpu() > 0, want_rb, want_lb)
        assert list(col.rewbuffer) == want_rb[-100:] and list(col.lenbuffer) == want_lb[-100:]
    assert len(want_lb) > 20


