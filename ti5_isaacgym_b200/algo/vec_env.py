"""The upper boundary of the hot path: what the reference's PPO runner expects of an environment.

The reference states it as an abstract base class (humanoid/algo/vec_env.py:6-31); what the runner actually touches is
listed in SURVEY.md 8(b) (dh_on_policy_runner.py:42-55, 68-73, 83, 101-105, 136; play.py:186-258).  Here the contract is
data — `CONTRACT` — plus a structural `VecEnv` protocol, so that `check_vec_env(env)` can tell a caller exactly which
part of the boundary an object misses (names, tensor dtypes, shapes), for this package's envs and for anybody's wrapper.
"""
from typing import Optional, Protocol, runtime_checkable

import torch

# name -> (kind, dtype or None, shape as a tuple of attribute names / ints or None)
CONTRACT = {
    "num_envs": ("int", None, None),
    "num_obs": ("int", None, None),
    "num_short_obs": ("int", None, None),
    "num_privileged_obs": ("int?", None, None),
    "num_actions": ("int", None, None),
    "max_episode_length": ("number", None, None),                 # a numpy float64 in the reference (lr:109)
    "obs_buf": ("tensor", torch.float32, ("num_envs", "num_obs")),
    "privileged_obs_buf": ("tensor?", torch.float32, ("num_envs", "num_privileged_obs")),
    "rew_buf": ("tensor", torch.float32, ("num_envs",)),
    "reset_buf": ("tensor", None, ("num_envs",)),                  # int64 before the first step, bool after (appendix A22)
    "episode_length_buf": ("tensor", torch.int64, ("num_envs",)),  # the runner writes it (runner :101)
    "extras": ("dict", None, None),
    "device": ("device", None, None),
}
METHODS = ("step", "reset", "get_observations", "get_privileged_observations")


@runtime_checkable
class VecEnv(Protocol):
    def step(self, actions: torch.Tensor): ...                     # -> obs, privileged obs | None, rewards, dones, infos

    def reset(self, *args): ...                                    # -> obs, privileged obs

    def get_observations(self) -> torch.Tensor: ...

    def get_privileged_observations(self) -> Optional[torch.Tensor]: ...


def check_vec_env(env):
    """Every way `env` falls short of the runner's contract, as a list of sentences (empty = drop-in)."""
    faults = [f"method {m}() is missing" for m in METHODS if not callable(getattr(env, m, None))]
    for name, (kind, dtype, shape) in CONTRACT.items():
        if not hasattr(env, name):
            faults.append(f"attribute {name} is missing")
            continue
        v = getattr(env, name)
        optional = kind.endswith("?")
        if v is None:
            if not optional:
                faults.append(f"{name} is None")
            continue
        base = kind.rstrip("?")
        if base == "int" and not isinstance(v, int):
            faults.append(f"{name} should be an int, is {type(v).__name__}")
        elif base == "number" and not hasattr(v, "__float__"):
            faults.append(f"{name} should be a number, is {type(v).__name__}")
        elif base == "dict" and not isinstance(v, dict):
            faults.append(f"{name} should be a dict, is {type(v).__name__}")
        elif base == "device" and not isinstance(v, (str, torch.device)):
            faults.append(f"{name} should be a device, is {type(v).__name__}")
        elif base == "tensor":
            if not torch.is_tensor(v):
                faults.append(f"{name} should be a tensor, is {type(v).__name__}")
                continue
            if dtype is not None and v.dtype != dtype:
                faults.append(f"{name} should be {dtype}, is {v.dtype}")
            want = tuple(getattr(env, s, None) if isinstance(s, str) else s for s in shape)
            if None not in want and tuple(v.shape) != want:
                faults.append(f"{name} should have shape {want}, has {tuple(v.shape)}")
    return faults
