"""The collection half of `DHOnPolicyRunner.learn` (dh_on_policy_runner.py:130-172) without its per-step host
round trips.

The reference loop does, every env step: `alg.act` -> `env.step` -> four `.to(device)` -> `alg.process_env_step`
(clone + bootstrap + ~10 `copy_` launches) -> `cur_reward_sum += ...`, `nonzero()`, two
`.cpu().numpy().tolist()` syncs.  Here the step's bookkeeping is the single `ti5_store_transition` launch of
`FrameLogRolloutStorage.store_step`; finished episodes are read back once per rollout.

`alg` is the reference's PPO object (or anything with `act(obs, critic_obs)`, `transition`, `gamma`,
`actor_critic.reset(dones)`, `compute_returns(last_critic_obs)`), unchanged.
"""
from collections import deque

import torch

from .rollout_storage import FrameLogRolloutStorage, install_frame_log_storage


class RolloutCollector:
    def __init__(self, env, alg, num_steps_per_env, group=None, buffer_len=100):
        self.env, self.alg, self.num_steps = env, alg, int(num_steps_per_env)
        if isinstance(getattr(alg, "storage", None), FrameLogRolloutStorage):
            self.storage = alg.storage
        else:
            self.storage = install_frame_log_storage(alg, env, group=group, num_transitions_per_env=self.num_steps)
        self.rewbuffer, self.lenbuffer = deque(maxlen=buffer_len), deque(maxlen=buffer_len)   # runner :119-121
        self.ep_infos = []
        self.obs = env.get_observations()
        critic = env.get_privileged_observations()
        self.critic_obs = critic if critic is not None else self.obs

    @torch.inference_mode()
    def collect(self):
        """runner :130-172: `num_steps` env steps into the storage, then returns / advantages.  Host syncs: one, at
        the end, for the finished-episode lists (the reference: two per step)."""
        env, alg, st = self.env, self.alg, self.storage
        st.clear()
        self.ep_infos.clear()
        obs, critic_obs = self.obs, self.critic_obs
        for _ in range(self.num_steps):
            actions = alg.act(obs, critic_obs)
            obs, privileged_obs, rewards, dones, infos = env.step(actions)
            critic_obs = privileged_obs if privileged_obs is not None else obs
            alg.process_env_step(rewards, dones, infos)                 # -> FrameLogRolloutStorage.store_step
            if "episode" in infos:
                self.ep_infos.append(infos["episode"])
        self.obs, self.critic_obs = obs, critic_obs
        alg.compute_returns(critic_obs)
        rew, length = st.finished_episodes()
        self.rewbuffer.extend(rew)
        self.lenbuffer.extend(length)
        return st
