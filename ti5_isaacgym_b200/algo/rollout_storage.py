"""`RolloutStorage` with the GAE scan on the GPU (reference: humanoid/algo/ppo/rollout_storage.py).

Only `compute_returns` (rs:97-119) is on the hot path: it becomes `ti5_gae_scan` + `ti5_gae_normalize`
(one thread per env walking T steps backwards, fp64 batch statistics).  The rest of the storage is
plain tensor bookkeeping that stays in torch, with the reference's field names, so `DHPPO`
(dh_ppo.py:67-110) can use this class unchanged; `patch_compute_returns` instead swaps the method
into the reference's own class.
"""
import ctypes

import torch

from .. import _lib


def gae_returns_(rewards, values, dones, last_values, returns, advantages, gamma, lam, scratch=None, group=None):
    """In-place GAE over (T, N, 1) tensors on a CUDA device.  With `group`, the (count, sum, sum of
    squares) of the advantages are all-reduced between the two kernels so every rank normalises
    with the global mean / unbiased std."""
    lib = _lib.load_library()
    T, N = rewards.shape[0], rewards.shape[1]
    for t in (rewards, values, dones, last_values, returns, advantages):
        if not (t.is_cuda and t.is_contiguous()):
            raise _lib.Ti5Error("ti5_gae needs contiguous CUDA tensors (there is no CPU fallback)")
    if dones.dtype not in (torch.uint8, torch.bool):
        raise _lib.Ti5Error("dones must be uint8 (rollout_storage.py:37)")
    if scratch is None:
        scratch = make_gae_scratch(N, rewards.device)
    stats, ticket = scratch
    st = ctypes.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    args = (ptr(rewards), ptr(values), ptr(dones), ptr(last_values), ptr(returns), ptr(advantages), T, N,
            float(gamma), float(lam), ptr(stats), ptr(ticket), st)
    if group is None:
        _lib.check(lib.ti5_gae(*args))
    else:
        import torch.distributed as dist
        _lib.check(lib.ti5_gae_scan(*args))
        dist.all_reduce(stats[:3], group=group)
        _lib.check(lib.ti5_gae_normalize(ptr(advantages), T, N, ptr(stats), st))
    return returns, advantages


def make_gae_scratch(num_envs, device):
    blocks = (num_envs + 127) // 128
    return (torch.zeros(4 + 2 * blocks, dtype=torch.float64, device=device),
            torch.zeros(1, dtype=torch.int32, device=device))


class RolloutStorage:
    class Transition:
        FIELDS = ("observations", "critic_observations", "actions", "rewards", "dones", "values", "actions_log_prob",
                  "action_mean", "action_sigma", "hidden_states", "next_proprio_obs")

        def __init__(self):
            for f in self.FIELDS:
                setattr(self, f, None)

        def clear(self):
            self.__init__()

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape,
                 num_single_obs=None, device="cpu", group=None):
        self.device = device
        self.obs_shape, self.privileged_obs_shape, self.actions_shape = obs_shape, privileged_obs_shape, actions_shape
        T, N = num_transitions_per_env, num_envs
        z = lambda *s: torch.zeros(T, N, *s, device=device)
        self.observations = z(*obs_shape)
        self.privileged_observations = z(*privileged_obs_shape) if privileged_obs_shape[0] is not None else None
        self.rewards, self.actions_log_prob, self.values, self.returns, self.advantages = z(1), z(1), z(1), z(1), z(1)
        self.actions, self.mu, self.sigma = z(*actions_shape), z(*actions_shape), z(*actions_shape)
        self.dones = z(1).byte()
        self.num_transitions_per_env, self.num_envs = T, N
        self.num_single_obs = num_single_obs
        if num_single_obs is not None:
            self.next_proprio_obs = z(num_single_obs)
        self.saved_hidden_states_a = self.saved_hidden_states_c = None
        self.step = 0
        self._group = group
        self._scratch = None

    def add_transitions(self, tr):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s = self.step
        self.observations[s].copy_(tr.observations)
        if self.privileged_observations is not None:
            self.privileged_observations[s].copy_(tr.critic_observations)
        self.actions[s].copy_(tr.actions)
        self.rewards[s].copy_(tr.rewards.view(-1, 1))
        self.dones[s].copy_(tr.dones.view(-1, 1))
        self.values[s].copy_(tr.values)
        self.actions_log_prob[s].copy_(tr.actions_log_prob.view(-1, 1))
        self.mu[s].copy_(tr.action_mean)
        self.sigma[s].copy_(tr.action_sigma)
        if self.num_single_obs is not None:
            self.next_proprio_obs[s].copy_(tr.next_proprio_obs)
        self.step += 1

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam):
        """rs:97-119 on the GPU; `self.returns` / `self.advantages` are updated in place."""
        if self._scratch is None:
            self._scratch = make_gae_scratch(self.num_envs, self.rewards.device)
        gae_returns_(self.rewards, self.values, self.dones, last_values.contiguous(), self.returns, self.advantages,
                     gamma, lam, self._scratch, self._group)

    def get_statistics(self):
        done = self.dones
        done[-1] = 1
        flat = done.permute(1, 0, 2).reshape(-1, 1)
        idx = torch.cat((flat.new_tensor([-1], dtype=torch.int64), flat.nonzero(as_tuple=False)[:, 0]))
        return (idx[1:] - idx[:-1]).float().mean(), self.rewards.mean()

    def mini_batch_generator(self, num_mini_batches, num_epochs=8):
        """rs:129-173: shuffled flat mini-batches (plain torch indexing)."""
        batch = self.num_envs * self.num_transitions_per_env
        mb = batch // num_mini_batches
        perm = torch.randperm(num_mini_batches * mb, requires_grad=False, device=self.device)
        flat = lambda t: t.flatten(0, 1)
        obs = flat(self.observations)
        crit = flat(self.privileged_observations) if self.privileged_observations is not None else obs
        cols = [flat(t) for t in (self.actions, self.values, self.advantages, self.returns, self.actions_log_prob,
                                  self.mu, self.sigma)]
        extra = [flat(self.next_proprio_obs), flat(self.rewards)] if self.num_single_obs is not None else []
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                idx = perm[i * mb:(i + 1) * mb]
                yield (*[t[idx] for t in extra], obs[idx], crit[idx], *[t[idx] for t in cols], (None, None), None)


def patch_compute_returns(storage_cls, group=None):
    """Swap the GPU GAE into an existing storage class (e.g. the reference's own RolloutStorage)."""
    def compute_returns(self, last_values, gamma, lam):
        if getattr(self, "_ti5_scratch", None) is None:
            self._ti5_scratch = make_gae_scratch(self.rewards.shape[1], self.rewards.device)
        # the reference REBINDS self.advantages each call (rs:118); keep that contract
        self.advantages = torch.empty_like(self.returns)
        gae_returns_(self.rewards, self.values, self.dones, last_values.contiguous(), self.returns, self.advantages,
                     gamma, lam, self._ti5_scratch, group)
    storage_cls.compute_returns = compute_returns
    return storage_cls
