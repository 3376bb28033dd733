"""Rollout-storage oracle against the reference's own RolloutStorage / DHPPO.process_env_step outputs
(tests/golden/rollout_T6_N24.npz, written by oracle/pin_rollout.py in the container that has the reference)."""
import os

import numpy as np
import pytest
import torch

from oracle import rollout_oracle as RO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_T6_N24.npz")
COLS = ("obs", "critic_obs", "actions", "values", "advantages", "returns", "actions_log_prob", "mu", "sigma")


def load_rollout_golden():
    z = np.load(GOLDEN)
    R = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in.") and z[k].ndim > 0}
    out = {k[4:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith("out.")}
    return R, float(z["in.gamma"]), out


def windows_from_frames(R):
    """Rebuild what the env returned at every step from the frame streams and the done flags (t1:474-479, 556-559)."""
    ho, hp = RO.Histories(list(R["pre_obs"])), RO.Histories(list(R["pre_priv"]))
    wins, cwins = [], []
    for t in range(R["rewards"].shape[0]):
        wins.append(ho.stack()); cwins.append(hp.stack())
        ids = R["dones"][t].nonzero(as_tuple=False).flatten()
        ho.clear(ids); hp.clear(ids)
        ho.append(R["obs_frames"][t]); hp.append(R["priv_frames"][t])
    return torch.stack(wins), torch.stack(cwins)


def test_rollout_oracle_reproduces_the_reference_storage():
    R, gamma, out = load_rollout_golden()
    R["windows"], R["critic_windows"] = windows_from_frames(R)
    T, N = R["rewards"].shape
    S = RO.Storage(T, N, R["windows"].shape[2], R["critic_windows"].shape[2], R["actions"].shape[2])
    cur_sum, cur_len, rb, lb = torch.zeros(N), torch.zeros(N), [], []
    for t in range(T):
        rew = RO.bootstrap_rewards(R["rewards"][t], R["values"][t], R["time_outs"][t], gamma)
        RO.add_transition(S, R["windows"][t], R["critic_windows"][t], R["actions"][t], rew, R["dones"][t], R["values"][t],
                          R["log_prob"][t], R["mean"][t], R["sigma"][t])
        RO.episode_bookkeeping(cur_sum, cur_len, R["rewards"][t], R["dones"][t], rb, lb)
    assert torch.equal(S.observations, out["observations"])
    assert torch.equal(S.privileged_observations, out["privileged_observations"])
    assert torch.equal(S.rewards, out["rewards"]) and torch.equal(S.dones, out["dones"])
    assert rb == out["rewbuffer"].tolist() and lb == out["lenbuffer"].tolist()
    S.returns.copy_(R["returns"]); S.advantages.copy_(R["advantages"])
    batches = list(RO.mini_batches(S, 4, 1, R["indices"]))
    assert len(batches) == 4
    for i, b in enumerate(batches):
        for c, v in zip(COLS, b):
            assert torch.equal(v, out[f"batch{i}.{c}"]), (i, c)


def test_overflow_raises_like_the_reference():
    S = RO.Storage(1, 2, 3, 3, 1)
    z = torch.zeros
    args = (z(2, 3), z(2, 3), z(2, 1), z(2), z(2, dtype=torch.bool), z(2, 1), z(2), z(2, 1), z(2, 1))
    RO.add_transition(S, *args)
    with pytest.raises(AssertionError, match="Rollout buffer overflow"):       # rs:60-61
        RO.add_transition(S, *args)
