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
        check_not_stale_ring_view(tr.observations, "observations")
        check_not_stale_ring_view(tr.critic_observations, "critic_observations")
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


class FrameLogRolloutStorage(RolloutStorage):
    """`RolloutStorage` without the (T, N, H*K) observation tensors (rs:30-33 — 2.4 GB at 8192 envs, written by a
    strided copy every step, rs:62-63, and gathered 32 times per iteration, rs:152-153).

    Consecutive observation windows of an env share H-1 of their H frames, so the env keeps each frame once
    (`LeggedRobot.enable_frame_log`) and `ti5_gather_minibatch` rebuilds the windows a mini-batch asks for; frames
    older than the env's last reset read as zero, as the reference's cleared deques do (t1:556-559).  One
    `ti5_store_transition` launch per step replaces the copies of `add_transitions` and, through `store_step`,
    the time-out bootstrap of `process_env_step` (dh_ppo.py:93-103) and the runner's episode bookkeeping
    (dh_on_policy_runner.py:156-168) — without its per-step host sync.

    The yielded batches are equal to the reference's for the same permutation (tests/test_rollout_gpu.py)."""

    def __init__(self, env, num_transitions_per_env, actions_shape=None, num_single_obs=None, device=None, group=None):
        if num_single_obs is not None:
            raise NotImplementedError("next_proprio_obs storage (rs:48-49) is unused by DHPPO (dh_ppo.py:70 passes None)")
        self.env = env
        self.device = device = torch.device(device if device is not None else env.device)
        if device.type != "cuda":
            raise _lib.Ti5Error("FrameLogRolloutStorage lives on the env's CUDA device (there is no CPU fallback)")
        T, N = int(num_transitions_per_env), env.num_envs
        self.obs_shape = [env.num_obs]
        self.privileged_obs_shape = [env.num_privileged_obs]
        self.actions_shape = list(actions_shape) if actions_shape is not None else [env.num_actions]
        A = self.actions_shape[0]
        z = lambda *sh, dtype=torch.float32: torch.zeros(*sh, dtype=dtype, device=device)
        self.rewards, self.actions_log_prob, self.values = z(T, N, 1), z(T, N, 1), z(T, N, 1)
        self.returns, self.advantages = z(T, N, 1), z(T, N, 1)
        self.actions, self.mu, self.sigma = z(T, N, A), z(T, N, A), z(T, N, A)
        self.dones = z(T, N, 1, dtype=torch.uint8)
        self.num_transitions_per_env, self.num_envs, self.num_single_obs = T, N, None
        self.saved_hidden_states_a = self.saved_hidden_states_c = None
        self.step = 0
        self._group, self._scratch = group, None
        # episode bookkeeping (runner :119-123, 156-168)
        self.cur_reward_sum, self.cur_episode_length = z(N), z(N)
        self._finished_rew, self._finished_len = z(T * N), z(T * N)
        self._n_finished = z(2, dtype=torch.int32)
        self._frame_row = z(T, dtype=torch.int32)
        env.enable_frame_log(T)
        self._lib = _lib.load_library()
        self._tr_struct = _lib.Ti5Transition()
        self._bind()

    def _bind(self):
        logs = self.env.frame_logs()
        ro = _lib.Ti5Rollout()
        ro.num_envs, ro.num_steps = self.num_envs, self.num_transitions_per_env
        ro.frame_stack, ro.c_frame_stack = logs.frame_stack, logs.c_frame_stack
        ro.num_single_obs, ro.priv_frame = logs.num_single_obs, logs.priv_frame
        ro.log_len, ro.num_actions = logs.log_len, self.actions_shape[0]
        self._logs = logs
        for name, t in (("frame_log", logs.frame_log), ("priv_log", logs.priv_log), ("valid_log", logs.valid_log),
                        ("frame_row", self._frame_row), ("actions", self.actions), ("mu", self.mu), ("sigma", self.sigma),
                        ("rewards", self.rewards), ("dones", self.dones), ("values", self.values),
                        ("actions_log_prob", self.actions_log_prob), ("returns", self.returns),
                        ("advantages", self.advantages), ("cur_reward_sum", self.cur_reward_sum),
                        ("cur_episode_length", self.cur_episode_length), ("finished_rew", self._finished_rew),
                        ("finished_len", self._finished_len), ("n_finished", self._n_finished)):
            setattr(ro, name, ctypes.c_void_p(t.data_ptr()))
        self._ro = ro
        self._ro_plain = _lib.Ti5Rollout.from_buffer_copy(ro)      # the same storage without the episode bookkeeping
        self._ro_plain.cur_reward_sum = None

    def _rollout_ref(self, episodes=True):
        if self._logs.frame_log is not getattr(self.env, "_frame_log", None):   # the env re-allocated its logs
            self._bind()
        return ctypes.byref(self._ro if episodes else self._ro_plain)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _f32(self, t, numel):
        if t.dtype is torch.float32 and t.is_cuda and t.is_contiguous() and t.numel() == numel and not t.requires_grad:
            return t                                      # the usual case: nothing to convert, no new tensor object
        t = t.detach()
        if t.dtype != torch.float32 or not t.is_cuda:
            t = t.to(device=self.device, dtype=torch.float32)
        t = t.contiguous()
        if t.numel() != numel:
            raise ValueError(f"expected {numel} elements, got a tensor of shape {tuple(t.shape)}")
        return t

    def _mask(self, t):
        if t.dtype in (torch.bool, torch.uint8) and t.is_cuda and t.is_contiguous():
            return t
        t = t.detach()
        if t.dtype not in (torch.bool, torch.uint8) or not t.is_cuda:
            t = (t.to(self.device) != 0)
        return t.contiguous().view(-1)

    def _store(self, tr, rewards, dones, time_outs, gamma, episodes):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")                        # rs:60-61
        N, A = self.num_envs, self.actions_shape[0]
        row = getattr(tr.observations, "ti5_frame_row", None)
        if row is None:
            # the runner's protocol: the stored observation is the one the policy acted on, i.e. the window
            # before the env step that has just completed (dh_ppo.py:88, runner :134-147)
            row = (self.env.frame_log_row - 1) % self._logs.log_len
        keep = [self._f32(tr.actions, N * A), self._f32(tr.action_mean, N * A), self._f32(tr.action_sigma, N * A),
                self._f32(tr.values, N), self._f32(tr.actions_log_prob, N), self._f32(rewards, N), self._mask(dones)]
        t = self._tr_struct                               # one struct, refilled: the call only borrows the pointers
        (t.actions, t.action_mean, t.action_sigma, t.values, t.actions_log_prob, t.rewards, t.dones) = [x.data_ptr() for x in keep]
        t.time_outs = None
        if time_outs is not None:
            keep.append(self._mask(time_outs))
            t.time_outs = keep[-1].data_ptr()
        _lib.check(self._lib.ti5_store_transition(self._rollout_ref(episodes), ctypes.byref(t), self.step, int(row),
                                                  float(gamma), self._stream()))
        self.step += 1

    def add_transitions(self, tr):
        """rs:59-74 with the reference's call protocol: `tr.rewards` already carries the time-out bootstrap."""
        self._store(tr, tr.rewards, tr.dones, None, 0.0, episodes=False)

    def store_step(self, tr, rewards, dones, time_outs, gamma):
        """dh_ppo.py:93-103 + rs:59-74 + runner :156-168 in one launch: `rewards` are the env's."""
        self._store(tr, rewards, dones, time_outs, gamma, episodes=True)

    def clear(self):
        self.step = 0
        self._n_finished.zero_()

    def finished_episodes(self):
        """(returns, lengths) of the episodes that ended during this rollout, in the order the runner's
        `rewbuffer.extend` / `lenbuffer.extend` see them (runner :162-164).  One device sync."""
        n = int(self._n_finished[self.step & 1].item())
        return self._finished_rew[:n].tolist(), self._finished_len[:n].tolist()

    # -- mini-batches ------------------------------------------------------------------------------------------
    def gather(self, idx, columns=("obs", "critic_obs", "actions", "values", "advantages", "returns",
                                   "actions_log_prob", "mu", "sigma"), sort=False):
        """Rows `idx` (flat t*N + e, rs:134-150) of the named columns as fresh tensors."""
        idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
        B, A = idx.numel(), self.actions_shape[0]
        widths = dict(obs=self.obs_shape[0], critic_obs=self.privileged_obs_shape[0], actions=A, values=1, advantages=1,
                      returns=1, actions_log_prob=1, mu=A, sigma=A)
        out, batch = {}, _lib.Ti5Batch()
        for c in columns:
            out[c] = torch.empty(B, widths[c], dtype=torch.float32, device=self.device)
            setattr(batch, c, ctypes.c_void_p(out[c].data_ptr()))
        order = None
        if sort and B > 1 and ("obs" in columns or "critic_obs" in columns):
            # produce the rows in (env, step) order: windows of one env overlap in all but a few frames.  Measured at
            # 8192 envs x 24 steps: the argsort costs more (0.12 ms) than the saved reads, so it is off by default
            N = self.num_envs
            order = torch.argsort((idx % N) * self.num_transitions_per_env + idx // N).to(torch.int32)
        _lib.check(self._lib.ti5_gather_minibatch(self._rollout_ref(), ctypes.c_void_p(idx.data_ptr()),
                                                  ctypes.c_void_p(order.data_ptr()) if order is not None else None, B,
                                                  ctypes.byref(batch), self._stream()))
        return out

    def mini_batch_generator(self, num_mini_batches, num_epochs=8, indices=None):
        """rs:129-173.  `indices` overrides the `torch.randperm` draw (parity tests)."""
        batch = self.num_envs * self.num_transitions_per_env
        mb = batch // num_mini_batches
        if indices is None:
            indices = torch.randperm(num_mini_batches * mb, requires_grad=False, device=self.device)
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                g = self.gather(indices[i * mb:(i + 1) * mb])
                yield (g["obs"], g["critic_obs"], g["actions"], g["values"], g["advantages"], g["returns"],
                       g["actions_log_prob"], g["mu"], g["sigma"], (None, None), None)

    def _all_rows(self, column):
        T, N = self.num_transitions_per_env, self.num_envs
        return self.gather(torch.arange(T * N, device=self.device), (column,))[column].view(T, N, -1)

    @property
    def observations(self):
        """(T, N, H*K) as the reference stores it — rebuilt on demand (debugging / parity checks)."""
        return self._all_rows("obs")

    @property
    def privileged_observations(self):
        return self._all_rows("critic_obs")


def check_not_stale_ring_view(t, what="observations"):
    """An env built with `materialize_obs=False` returns views into its history rings that the NEXT `step()` overwrites
    (oldest slot reused, re-spawned envs cleared).  The reference's runner stores the observation one step after the
    policy saw it (dh_ppo.py:88 -> rs:62): copying such a view then would silently store the wrong window.  The env
    tags its views with the step they belong to; a copy-late storage calls this and fails loudly instead."""
    tag = getattr(t, "ti5_ring_view", None)
    if tag is not None and tag[0][0] != tag[1]:
        raise _lib.Ti5Error(f"{what}: a view into the env's history ring taken at step {tag[1]}, but the env has stepped "
                            f"to {tag[0][0]} since — the window it showed is gone.  Build the env with materialize_obs=True "
                            "(the default), or use FrameLogRolloutStorage (task_registry.make_alg_runner does).")


def install_frame_log_storage(alg, env, group=None, num_transitions_per_env=None):
    """Swap a `FrameLogRolloutStorage` into a PPO object built by the reference's runner (`alg.storage`,
    dh_ppo.py:67-70) and fuse its `process_env_step` (dh_ppo.py:93-103) into the single store launch."""
    old = getattr(alg, "storage", None)
    T = num_transitions_per_env if old is None else old.num_transitions_per_env
    shape = None if old is None else old.actions_shape
    alg.storage = None
    del old                                   # release the (T, N, H*K) tensors before allocating anything else
    storage = FrameLogRolloutStorage(env, T, shape, device=env.device, group=group)
    alg.storage = storage

    def process_env_step(rewards, dones, infos):
        storage.store_step(alg.transition, rewards, dones, infos.get("time_outs"), alg.gamma)
        alg.transition.clear()
        alg.actor_critic.reset(dones)
    alg.process_env_step = process_env_step
    return storage


def patch_compute_returns(storage_cls, group=None):
    """Swap the GPU GAE into an existing storage class (e.g. the reference's own RolloutStorage)."""
    add_plain = storage_cls.add_transitions

    def add_transitions(self, transition):
        check_not_stale_ring_view(transition.observations, "observations")
        check_not_stale_ring_view(getattr(transition, "critic_observations", None), "critic_observations")
        return add_plain(self, transition)
    storage_cls.add_transitions = add_transitions

    def compute_returns(self, last_values, gamma, lam):
        if getattr(self, "_ti5_scratch", None) is None:
            self._ti5_scratch = make_gae_scratch(self.rewards.shape[1], self.rewards.device)
        # the reference REBINDS self.advantages each call (rs:118); keep that contract
        self.advantages = torch.empty_like(self.returns)
        gae_returns_(self.rewards, self.values, self.dones, last_values.contiguous(), self.returns, self.advantages,
                     gamma, lam, self._ti5_scratch, group)
    storage_cls.compute_returns = compute_returns
    return storage_cls
