"""Pin oracle/rollout_oracle.py against the reference's own classes, and write the golden fixture.

Runs ONLY in the build container (needs /root/reference).  The reference's unmodified `RolloutStorage`
(rollout_storage.py) is filled through the unmodified `DHPPO.process_env_step` (dh_ppo.py:93-103; the PPO object
is built without its networks — only `transition`, `gamma`, `storage`, `device` and an `actor_critic.reset`
stub are touched by that method), the runner's bookkeeping lines (dh_on_policy_runner.py:156-168) are executed
next to it, and `mini_batch_generator` is driven with a fixed permutation (torch.randperm patched for the call).
The oracle must agree BIT-FOR-BIT on every stored tensor, both episode lists and every mini-batch.

    python oracle/pin_rollout.py            # check only
    python oracle/pin_rollout.py --write    # check + regenerate tests/golden/rollout_T6_N24.npz
"""
import argparse
import os
import sys
from collections import deque

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(1, "/root/reference")

from oracle import rollout_oracle as RO     # noqa: E402

BATCH_COLS = ("obs", "critic_obs", "actions", "values", "advantages", "returns", "actions_log_prob", "mu", "sigma")


def reference_rollout(R, gamma, num_mini_batches, num_epochs, indices):
    import humanoid.envs  # noqa: F401  (import order: envs before algo, as train.py does)
    from humanoid.algo.ppo.dh_ppo import DHPPO
    from humanoid.algo.ppo.rollout_storage import RolloutStorage
    T, N = R["rewards"].shape
    alg = object.__new__(DHPPO)
    alg.device, alg.gamma = "cpu", gamma
    alg.transition = RolloutStorage.Transition()
    alg.actor_critic = type("Stub", (), {"reset": staticmethod(lambda dones: None)})()
    alg.storage = RolloutStorage(N, T, [R["windows"].shape[2]], [R["critic_windows"].shape[2]], [R["actions"].shape[2]],
                                 None, "cpu")
    cur_reward_sum, cur_episode_length = torch.zeros(N), torch.zeros(N)
    rewbuffer, lenbuffer = deque(maxlen=10 ** 9), deque(maxlen=10 ** 9)
    for t in range(T):
        tr = alg.transition                                    # what DHPPO.act leaves behind (dh_ppo.py:76-91)
        tr.actions, tr.values, tr.actions_log_prob = R["actions"][t], R["values"][t], R["log_prob"][t]
        tr.action_mean, tr.action_sigma = R["mean"][t], R["sigma"][t]
        tr.observations, tr.critic_observations = R["windows"][t], R["critic_windows"][t]
        rewards, dones, infos = R["rewards"][t].clone(), R["dones"][t], {"time_outs": R["time_outs"][t]}
        alg.process_env_step(rewards, dones, infos)
        # dh_on_policy_runner.py:156-168, verbatim semantics
        cur_reward_sum += rewards
        cur_episode_length += 1
        new_ids = (dones > 0).nonzero(as_tuple=False)
        rewbuffer.extend(cur_reward_sum[new_ids][:, 0].cpu().numpy().tolist())
        lenbuffer.extend(cur_episode_length[new_ids][:, 0].cpu().numpy().tolist())
        cur_reward_sum[new_ids] = 0
        cur_episode_length[new_ids] = 0
    S = alg.storage
    # give returns / advantages recognisable content (compute_returns is pinned separately)
    S.returns.copy_(torch.arange(T * N, dtype=torch.float32).view(T, N, 1) * 0.5)
    S.advantages = S.returns * -2.0 + 1.0
    real = torch.randperm
    torch.randperm = lambda n, **kw: indices[:n].clone()
    try:
        batches = [b[:9] for b in S.mini_batch_generator(num_mini_batches, num_epochs)]
    finally:
        torch.randperm = real
    return S, list(rewbuffer), list(lenbuffer), batches


def check(T, N, H, CH, K, P, A, seed, gamma=0.994, num_mini_batches=4, num_epochs=2):
    R = RO.synthetic_rollout(T, N, H, CH, K, P, A, seed)
    mb = (T * N) // num_mini_batches
    indices = torch.randperm(num_mini_batches * mb, generator=torch.Generator().manual_seed(seed + 1))
    Sr, rb, lb, br = reference_rollout(R, gamma, num_mini_batches, num_epochs, indices)
    So = RO.Storage(T, N, H * K, CH * P, A)
    cur_sum, cur_len, rbo, lbo = torch.zeros(N), torch.zeros(N), [], []
    for t in range(T):
        rew = RO.bootstrap_rewards(R["rewards"][t], R["values"][t], R["time_outs"][t], gamma)
        RO.add_transition(So, R["windows"][t], R["critic_windows"][t], R["actions"][t], rew, R["dones"][t],
                          R["values"][t], R["log_prob"][t], R["mean"][t], R["sigma"][t])
        RO.episode_bookkeeping(cur_sum, cur_len, R["rewards"][t], R["dones"][t], rbo, lbo)
    So.returns.copy_(Sr.returns); So.advantages.copy_(Sr.advantages)
    for name in ("observations", "privileged_observations", "actions", "rewards", "dones", "values", "actions_log_prob", "mu", "sigma"):
        assert torch.equal(getattr(Sr, name), getattr(So, name)), name
    assert rb == rbo and lb == lbo, "episode lists"
    bo = list(RO.mini_batches(So, num_mini_batches, num_epochs, indices))
    assert len(bo) == len(br)
    for x, y in zip(br, bo):
        for c, u, v in zip(BATCH_COLS, x, y):
            assert torch.equal(u, v), c
    print(f"rollout T={T} N={N} H={H} seed={seed}: oracle == reference  ({len(rb)} finished episodes, {len(br)} batches)")
    return R, indices, Sr, rb, lb, br


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", action="store_true")
    args = ap.parse_args()
    check(24, 64, 66, 3, 47, 73, 12, 5)
    check(5, 33, 4, 3, 47, 260, 12, 9, num_mini_batches=3)
    R, indices, S, rb, lb, batches = check(6, 24, 5, 3, 47, 73, 12, 3, num_mini_batches=4, num_epochs=1)
    if args.write:
        out = {f"in.{k}": v.numpy() for k, v in R.items() if k not in ("windows", "critic_windows")}
        out["in.indices"] = indices.numpy()
        out["in.gamma"] = np.float64(0.994)
        out["in.returns"], out["in.advantages"] = S.returns.numpy(), S.advantages.numpy()
        out["out.rewards"], out["out.dones"] = S.rewards.numpy(), S.dones.numpy()
        out["out.observations"], out["out.privileged_observations"] = S.observations.numpy(), S.privileged_observations.numpy()
        out["out.rewbuffer"], out["out.lenbuffer"] = np.asarray(rb, np.float64), np.asarray(lb, np.float64)
        for i, b in enumerate(batches):
            for c, v in zip(BATCH_COLS, b):
                out[f"out.batch{i}.{c}"] = v.numpy()
        path = os.path.join(ROOT, "tests", "golden", "rollout_T6_N24.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")
