"""Record the call signatures of the reference's classes on both sides of the hot path (what calls the env, what the
env's storage is called with) into tests/golden/reference_signatures.json.  TEST INFRASTRUCTURE; runs only where the
reference tree is mounted.  The GPU box has no reference: there the stand-in PPO objects, this package's env / storage
classes and the collector are checked against the recorded signatures.

    python oracle/pin_signatures.py            # print
    python oracle/pin_signatures.py --write    # regenerate the fixture
"""
import inspect
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _sig(fn):
    return [p.name if p.default is inspect.Parameter.empty else f"{p.name}={p.default!r}"
            for p in inspect.signature(fn).parameters.values()]


def collect():
    from oracle.reference_driver import import_reference
    import_reference()
    from humanoid.algo.ppo.dh_ppo import DHPPO
    from humanoid.algo.ppo.rollout_storage import RolloutStorage
    from humanoid.algo.ppo.dh_on_policy_runner import DHOnPolicyRunner
    from humanoid.algo.vec_env import VecEnv
    from humanoid.envs.base.legged_robot import LeggedRobot
    from humanoid.envs.t1.t1_dh_stand_env import T1DHStandEnv
    from humanoid.utils.task_registry import TaskRegistry
    out = {}
    for cls, names in ((DHPPO, ("act", "process_env_step", "compute_returns", "init_storage", "update")),
                       (RolloutStorage, ("__init__", "add_transitions", "clear", "compute_returns", "mini_batch_generator")),
                       (DHOnPolicyRunner, ("__init__", "learn", "save", "load", "get_inference_policy")),
                       (VecEnv, ("step", "reset", "get_observations", "get_privileged_observations")),
                       (LeggedRobot, ("__init__", "step", "reset", "reset_idx", "post_physics_step", "_compute_torques",
                                      "_get_heights", "_refresh_actor_dof_props", "check_termination", "compute_reward")),
                       (T1DHStandEnv, ("__init__", "step", "reset_idx", "compute_observations", "compute_ref_state")),
                       (TaskRegistry, ("register", "get_task_class", "get_cfgs", "make_env", "make_alg_runner"))):
        for n in names:
            out[f"{cls.__name__}.{n}"] = _sig(getattr(cls, n))
    out["RolloutStorage.Transition.fields"] = sorted(vars(RolloutStorage.Transition()).keys())
    out["VecEnv.annotations"] = sorted(getattr(VecEnv, "__annotations__", {}).keys())
    return out


if __name__ == "__main__":
    sigs = collect()
    if "--write" in sys.argv:
        with open(os.path.join(ROOT, "tests", "golden", "reference_signatures.json"), "w") as f:
            json.dump(sigs, f, indent=1, sort_keys=True)
    print(json.dumps(sigs, indent=1, sort_keys=True))
