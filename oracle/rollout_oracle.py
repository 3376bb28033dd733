"""CPU restatement of the rollout-storage path (TEST INFRASTRUCTURE — never imported by the product).

Follows, line by line:
  rs   = humanoid/algo/ppo/rollout_storage.py   (add_transitions :59-74, mini_batch_generator :129-173)
  ppo  = humanoid/algo/ppo/dh_ppo.py            (process_env_step :93-103)
  run  = humanoid/algo/ppo/dh_on_policy_runner.py (episode bookkeeping :156-168)
  t1   = humanoid/envs/t1/t1_dh_stand_env.py    (history deques: append :474-475, clear on reset :556-559, stack :477-479)

Pinned against the reference's own `RolloutStorage` / `DHPPO.process_env_step` by oracle/pin_rollout.py
(bit-equal on the committed fixture tests/golden/rollout_T6_N24.npz and on random cases).
"""
import torch


class Storage:
    """The tensors of rs:29-44 (no hidden states, no next_proprio_obs: DHPPO passes None, ppo:70)."""

    def __init__(self, T, N, obs_dim, priv_dim, A):
        z = torch.zeros
        self.observations, self.privileged_observations = z(T, N, obs_dim), z(T, N, priv_dim)
        self.rewards, self.actions_log_prob, self.values, self.returns, self.advantages = (z(T, N, 1) for _ in range(5))
        self.actions, self.mu, self.sigma = z(T, N, A), z(T, N, A), z(T, N, A)
        self.dones = z(T, N, 1).byte()
        self.T, self.N, self.step = T, N, 0


def bootstrap_rewards(rewards, values, time_outs, gamma):
    """ppo:95-98: rewards.clone(); rewards += gamma * squeeze(values * time_outs.unsqueeze(1), 1)."""
    out = rewards.clone()
    if time_outs is not None:
        out += gamma * torch.squeeze(values * time_outs.unsqueeze(1), 1)
    return out


def add_transition(S, obs, critic_obs, actions, rewards, dones, values, log_prob, mean, sigma):
    """rs:59-74."""
    if S.step >= S.T:
        raise AssertionError("Rollout buffer overflow")
    s = S.step
    S.observations[s].copy_(obs)
    S.privileged_observations[s].copy_(critic_obs)
    S.actions[s].copy_(actions)
    S.rewards[s].copy_(rewards.view(-1, 1))
    S.dones[s].copy_(dones.view(-1, 1))
    S.values[s].copy_(values)
    S.actions_log_prob[s].copy_(log_prob.view(-1, 1))
    S.mu[s].copy_(mean)
    S.sigma[s].copy_(sigma)
    S.step += 1


def episode_bookkeeping(cur_reward_sum, cur_episode_length, rewards, dones, rewbuffer, lenbuffer):
    """run:156-168 (in place; the buffers are plain lists here, deque(maxlen=100) in the runner)."""
    cur_reward_sum += rewards
    cur_episode_length += 1
    new_ids = (dones > 0).nonzero(as_tuple=False)
    rewbuffer.extend(cur_reward_sum[new_ids][:, 0].cpu().numpy().tolist())
    lenbuffer.extend(cur_episode_length[new_ids][:, 0].cpu().numpy().tolist())
    cur_reward_sum[new_ids] = 0
    cur_episode_length[new_ids] = 0


def mini_batches(S, num_mini_batches, num_epochs, indices):
    """rs:129-173 with the permutation given (the reference draws it with torch.randperm, rs:132)."""
    mb = (S.N * S.T) // num_mini_batches
    cols = [t.flatten(0, 1) for t in (S.observations, S.privileged_observations, S.actions, S.values, S.advantages,
                                      S.returns, S.actions_log_prob, S.mu, S.sigma)]
    for _ in range(num_epochs):
        for i in range(num_mini_batches):
            idx = indices[i * mb:(i + 1) * mb]
            yield tuple(c[idx] for c in cols)


class Histories:
    """The env's observation deques as the storage sees them: `append` (t1:474-475), `clear` for re-spawned envs
    (t1:556-559, before the append of the same step), `stack` oldest -> newest (t1:477-479)."""

    def __init__(self, frames):            # frames: list of (N, W) tensors, oldest first
        self.frames = [f.clone() for f in frames]

    def clear(self, env_ids):
        for f in self.frames:
            f[env_ids] *= 0

    def append(self, frame):
        self.frames = self.frames[1:] + [frame.clone()]

    def stack(self):
        N = self.frames[0].shape[0]
        return torch.stack(self.frames, dim=1).reshape(N, -1)


def synthetic_rollout(T, N, H, CH, K, P, A, seed, done_p=0.15, timeout_p=0.5):
    """Seeded inputs of a rollout with resets: frame streams, the windows the env would return, the policy
    outputs and the env outputs per step.  Window t is what the policy acts on at step t; `dones[t]` re-spawns
    envs inside env step t, i.e. before frame t+1 is appended."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    pre_obs, pre_priv = [rn(N, K) for _ in range(H)], [rn(N, P) for _ in range(CH)]
    ho, hp = Histories(pre_obs), Histories(pre_priv)
    out = dict(pre_obs=torch.stack(pre_obs), pre_priv=torch.stack(pre_priv), obs_frames=[], priv_frames=[], windows=[],
               critic_windows=[], actions=[], values=[], log_prob=[], mean=[], sigma=[], rewards=[], dones=[], time_outs=[])
    for t in range(T):
        out["windows"].append(ho.stack())
        out["critic_windows"].append(hp.stack())
        out["actions"].append(rn(N, A)); out["values"].append(rn(N, 1)); out["log_prob"].append(rn(N))
        out["mean"].append(rn(N, A)); out["sigma"].append(rn(N, A).abs() + 0.1)
        out["rewards"].append(rn(N))
        dones = torch.rand(N, generator=g) < done_p
        out["dones"].append(dones)
        out["time_outs"].append(dones & (torch.rand(N, generator=g) < timeout_p))
        ids = dones.nonzero(as_tuple=False).flatten()
        ho.clear(ids); hp.clear(ids)
        fo, fp = rn(N, K), rn(N, P)
        ho.append(fo); hp.append(fp)
        out["obs_frames"].append(fo); out["priv_frames"].append(fp)
    return {k: (torch.stack(v) if isinstance(v, list) else v) for k, v in out.items()}


def run_rollout(R, gamma, num_mini_batches, num_epochs, indices):
    """The reference's collection loop on the inputs of `synthetic_rollout`: ppo:93-103 -> rs:59-74 per step,
    run:156-168 bookkeeping, then the mini-batches.  Returns (storage, rewbuffer, lenbuffer, batches)."""
    T, N = R["rewards"].shape
    S = Storage(T, N, R["windows"].shape[2], R["critic_windows"].shape[2], R["actions"].shape[2])
    cur_sum, cur_len, rewbuf, lenbuf = torch.zeros(N), torch.zeros(N), [], []
    for t in range(T):
        rew = bootstrap_rewards(R["rewards"][t], R["values"][t], R["time_outs"][t], gamma)
        add_transition(S, R["windows"][t], R["critic_windows"][t], R["actions"][t], rew, R["dones"][t], R["values"][t],
                       R["log_prob"][t], R["mean"][t], R["sigma"][t])
        episode_bookkeeping(cur_sum, cur_len, R["rewards"][t], R["dones"][t], rewbuf, lenbuf)
    return S, rewbuf, lenbuf, list(mini_batches(S, num_mini_batches, num_epochs, indices))
