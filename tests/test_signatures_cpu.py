"""The call signatures on both sides of the hot path equal the reference's (tests/golden/reference_signatures.json, written
by oracle/pin_signatures.py from the unmodified reference tree): this package's host mirror is a drop-in at the level
the reference's callers bind to — same parameter names, order and defaults (extra keyword-only options may follow)."""
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = json.load(open(os.path.join(HERE, "golden", "reference_signatures.json")))


def sig(fn):
    return [("**" + p.name if p.kind is inspect.Parameter.VAR_KEYWORD else p.name) if p.default is inspect.Parameter.empty
            else f"{p.name}={p.default!r}" for p in inspect.signature(fn).parameters.values()]


def leading(ours, theirs):
    """`ours` starts with the reference's parameters (names and defaults); whatever follows must have defaults."""
    names = lambda s: [x.split("=")[0] for x in s]
    assert names(ours)[:len(theirs)] == names(theirs), (ours, theirs)
    for extra in ours[len(theirs):]:
        assert "=" in extra or extra.startswith("**"), f"extra parameter {extra} needs a default"


@pytest.mark.parametrize("name", ["LeggedRobot.__init__", "LeggedRobot.step", "LeggedRobot.reset", "LeggedRobot.reset_idx",
                                  "LeggedRobot.post_physics_step", "LeggedRobot._get_heights",
                                  "LeggedRobot._refresh_actor_dof_props", "T1DHStandEnv.step",
                                  "T1DHStandEnv.compute_observations"])
def test_env_methods(name):
    from ti5_isaacgym_b200.envs.base.legged_robot import LeggedRobot
    from ti5_isaacgym_b200.envs.t1.t1_dh_stand_env import T1DHStandEnv
    cls, meth = name.split(".")
    ours = sig(getattr({"LeggedRobot": LeggedRobot, "T1DHStandEnv": T1DHStandEnv}[cls], meth))
    leading(ours, REF[name])


@pytest.mark.parametrize("name", ["RolloutStorage.__init__", "RolloutStorage.add_transitions", "RolloutStorage.clear",
                                  "RolloutStorage.compute_returns", "RolloutStorage.mini_batch_generator"])
def test_storage_methods(name):
    from ti5_isaacgym_b200.algo.rollout_storage import RolloutStorage
    ours = sig(getattr(RolloutStorage, name.split(".")[1]))
    theirs = [x for x in REF[name]]
    names = lambda s: [x.split("=")[0] for x in s]
    # the reference calls its transition parameter `transition`; positional use only (dh_ppo.py:99)
    if name.endswith("add_transitions"):
        assert len(ours) == len(theirs) == 2
        return
    assert names(ours)[:len(theirs)] == names(theirs), (ours, theirs)


def test_transition_fields_and_registry():
    from ti5_isaacgym_b200.algo.rollout_storage import RolloutStorage
    from ti5_isaacgym_b200.utils.task_registry import TaskRegistry
    assert sorted(vars(RolloutStorage.Transition()).keys()) == REF["RolloutStorage.Transition.fields"]
    for meth in ("register", "get_task_class", "get_cfgs", "make_env", "make_alg_runner"):
        leading(sig(getattr(TaskRegistry, meth)), REF[f"TaskRegistry.{meth}"])


def test_vec_env_contract_lists_what_the_reference_abc_declares():
    from ti5_isaacgym_b200.algo.vec_env import CONTRACT, METHODS
    assert set(METHODS) == {"step", "reset", "get_observations", "get_privileged_observations"}
    declared = set(REF["VecEnv.annotations"])          # attributes annotated on the reference's ABC (vec_env.py:7-17)
    assert declared <= set(CONTRACT) | {"num_privileged_obs"}, declared - set(CONTRACT)
