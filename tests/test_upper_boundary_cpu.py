"""Upper boundary: the UNMODIFIED `DHOnPolicyRunner` of the reference (dh_on_policy_runner.py:21-201,
with its own DHPPO, ActorCriticDH and RolloutStorage) runs one full learning iteration — 24 env steps, GAE, the PPO
update — against an environment object that exposes NOTHING but `algo.vec_env.CONTRACT` + `METHODS` (and the four cfg
fields the runner reads), with the oracle doing the step math behind it on the CPU.  Every attribute the runner touches is
recorded: the contract list is thereby pinned to what the real caller needs, and `check_vec_env` is shown to be the
right gate for the CUDA env.  Needs the reference tree (build container only)."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

from oracle import t1_oracle as O

pytestmark = pytest.mark.reference


class ContractEnv:
    """A VecEnv that answers only to the names of the contract; everything else raises AttributeError."""

    def __init__(self, N, seed=0):
        from ti5_isaacgym_b200.algo.vec_env import CONTRACT, METHODS
        from ti5_isaacgym_b200.envs import DHT1StandCfg
        from ti5_isaacgym_b200.envs.t1.t1_robot import robot_constants
        from ti5_isaacgym_b200.sim.synthetic import alloc_sim_tensors
        cfg = DHT1StandCfg()
        cfg.env.num_envs = N
        cfg.terrain.mesh_type = "plane"
        d = object.__getattribute__(self, "__dict__")
        d["_allowed"] = set(CONTRACT) | set(METHODS) | {"cfg", "num_single_obs"}
        d["_touched"] = set()
        d["_cfg_touched"] = set()
        C = O.make_consts(cfg, cfg.sim.dt, robot_constants(cfg))
        S = O.new_state(C, N)
        gen = torch.Generator().manual_seed(seed)
        S.lag_timestep[:] = torch.randint(0, 31, (N,), generator=gen)
        S.gait_time[:, 1], S.gait_time[:, 2] = 900, 1500
        S.randomized_p_gains[:], S.randomized_d_gains[:] = C.p_gains, C.d_gains
        d["_o"] = SimpleNamespace(C=C, S=S, gen=gen, sim=alloc_sim_tensors(N, "cpu"), cfg=cfg)
        # the contract's data
        d["num_envs"], d["num_obs"], d["num_short_obs"] = N, cfg.env.num_observations, cfg.env.short_frame_stack * cfg.env.num_single_obs
        d["num_single_obs"] = cfg.env.num_single_obs
        d["num_privileged_obs"], d["num_actions"] = cfg.env.num_privileged_obs, cfg.env.num_actions
        d["max_episode_length"], d["device"], d["extras"] = C.max_episode_length, "cpu", {}
        d["obs_buf"], d["privileged_obs_buf"] = S.obs_buf, S.privileged_obs_buf
        d["rew_buf"], d["reset_buf"], d["episode_length_buf"] = S.rew_buf, S.reset_buf, S.episode_length_buf

    def __getattribute__(self, name):
        if name.startswith("__") or name.startswith("_"):
            return object.__getattribute__(self, name)
        d = object.__getattribute__(self, "__dict__")
        if name not in d["_allowed"]:
            raise AttributeError(f"the runner touched env.{name}, which is not in the VecEnv contract")
        d["_touched"].add(name)
        if name == "cfg":
            return _Recorder(d["_o"].cfg, "cfg", d["_cfg_touched"])
        return object.__getattribute__(self, name)

    def __setattr__(self, name, value):
        d = object.__getattribute__(self, "__dict__")
        if name not in d["_allowed"]:
            raise AttributeError(f"the runner assigned env.{name}, which is not in the VecEnv contract")
        d["_touched"].add(name + "=")
        if name == "episode_length_buf":           # the runner REBINDS it (runner :101)
            d["_o"].S.episode_length_buf = value
        d[name] = value

    # -- methods of the contract
    def _one_step(self, actions):
        from ti5_isaacgym_b200.sim.synthetic import fill_synthetic_state
        o = self._o
        fill_synthetic_state(o.sim, o.S.env_origins, o.gen, base_contact_rate=0.03)
        pools = O.draw_pools(o.C, self.__dict__["num_envs"], o.gen)
        obs, priv, rew, reset, extras = O.step(o.C, o.S, o.sim, actions, pools)
        d = self.__dict__
        d["obs_buf"], d["privileged_obs_buf"], d["rew_buf"], d["reset_buf"] = obs, priv, rew, reset
        d["episode_length_buf"], d["extras"] = o.S.episode_length_buf, extras
        return obs, priv, rew, reset, extras

    def step(self, actions):
        return self._one_step(actions)

    def reset(self):
        d = self.__dict__
        obs, priv, _, _, _ = self._one_step(torch.zeros(d["num_envs"], d["num_actions"]))
        return obs, priv

    def get_observations(self):
        return self.__dict__["obs_buf"]

    def get_privileged_observations(self):
        return self.__dict__["privileged_obs_buf"]


class _Recorder:
    """Records dotted attribute paths read through it."""

    def __init__(self, obj, path, log):
        object.__setattr__(self, "_s", (obj, path, log))

    def __getattr__(self, name):
        obj, path, log = object.__getattribute__(self, "_s")
        v = getattr(obj, name)
        p = f"{path}.{name}"
        if isinstance(v, type) or hasattr(v, "__dict__") and not callable(v):
            return _Recorder(v, p, log)
        log.add(p)
        return v


def test_unmodified_runner_learns_one_iteration_on_a_contract_only_env(tmp_path):
    from oracle.reference_driver import import_reference
    import_reference()
    from humanoid.algo import DHOnPolicyRunner
    from humanoid.envs import DHT1StandCfgPPO
    from humanoid.utils.helpers import class_to_dict as ref_class_to_dict
    from ti5_isaacgym_b200.algo.vec_env import CONTRACT, METHODS, check_vec_env
    torch.manual_seed(0)
    N = 8
    env = ContractEnv(N)
    assert check_vec_env(env) == [], "the contract env must pass the gate it defines"
    env.__dict__["_touched"].clear()
    train_cfg = ref_class_to_dict(DHT1StandCfgPPO())
    runner = DHOnPolicyRunner(env, train_cfg, log_dir=str(tmp_path), device="cpu")     # with its TensorBoard logging
    before = [p.detach().clone() for p in runner.alg.actor_critic.parameters()]
    runner.learn(1, init_at_random_ep_len=True)
    assert any(f.startswith("model_") for f in os.listdir(tmp_path)), "runner.save wrote its checkpoint"
    after = list(runner.alg.actor_critic.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(after, before)), "the PPO update must have stepped the policy"
    st = runner.alg.storage
    assert st.observations.shape == (24, N, 66 * 47) and st.privileged_observations.shape == (24, N, 3 * 73)
    assert torch.isfinite(st.returns).all() and torch.isfinite(st.advantages).all() and st.step == 0
    touched = {t.rstrip("=") for t in env.__dict__["_touched"]}
    # everything the runner touched is in the contract ...
    assert touched <= set(CONTRACT) | set(METHODS) | {"cfg", "num_single_obs"}
    # ... and the contract lists nothing the runner does not need, apart from what play.py / the VecEnv ABC add
    unused = (set(CONTRACT) | set(METHODS)) - touched
    assert unused <= {"obs_buf", "privileged_obs_buf", "rew_buf", "reset_buf", "extras", "device"}, unused
    assert "episode_length_buf=" in env.__dict__["_touched"], "runner :101 rebinds episode_length_buf"
    assert env.__dict__["_cfg_touched"] <= {"cfg.terrain.measure_heights", "cfg.env.c_frame_stack",
                                            "cfg.env.single_num_privileged_obs", "cfg.terrain.num_height"}


def test_reference_signatures_fixture_is_current():
    """tests/golden/reference_signatures.json (what the GPU-side stand-ins are checked against) still matches the
    reference tree."""
    import json
    from oracle.pin_signatures import collect
    here = os.path.dirname(os.path.abspath(__file__))
    assert json.load(open(os.path.join(here, "golden", "reference_signatures.json"))) == collect()
