"""Host logic of `utils.helpers` / `utils.task_registry` that needs no GPU: the checkpoint look-up of the reference
(humanoid/utils/helpers.py:94-123) and the resume branch of `make_alg_runner` (humanoid/utils/task_registry.py:137-143)."""
import os
import sys
import types

import pytest

from ti5_isaacgym_b200.utils.helpers import get_load_path, update_cfg_from_args
from ti5_isaacgym_b200.utils.task_registry import TaskRegistry, default_args


def _tree(root):
    for run, models in (("2024-01-01_10-00-00ti5", [100, 900]), ("2024-02-01_10-00-00ti5", [100, 900, 2000]), ("exported", [])):
        os.makedirs(os.path.join(root, run))
        for it in models:
            open(os.path.join(root, run, f"model_{it}.pt"), "w").close()
    open(os.path.join(root, "2024-02-01_10-00-00ti5", "events.out"), "w").close()     # not a model file


def test_get_load_path_picks_the_newest_run_and_model(tmp_path):
    root = str(tmp_path)
    _tree(root)
    # newest run by name ("exported" skipped), newest model by zero-padded name (2000 after 900, not before)
    assert get_load_path(root) == os.path.join(root, "2024-02-01_10-00-00ti5", "model_2000.pt")
    assert get_load_path(root, load_run="2024-01-01_10-00-00ti5") == os.path.join(root, "2024-01-01_10-00-00ti5", "model_900.pt")
    assert get_load_path(root, checkpoint=100) == os.path.join(root, "2024-02-01_10-00-00ti5", "model_100.pt")
    assert get_load_path(root, load_run="2024-01-01_10-00-00ti5", checkpoint=100).endswith(
        os.path.join("2024-01-01_10-00-00ti5", "model_100.pt"))


def test_get_load_path_without_runs_raises_like_the_reference(tmp_path):
    with pytest.raises(ValueError, match="No runs in this directory"):
        get_load_path(str(tmp_path))
    with pytest.raises(ValueError, match="No runs in this directory"):
        get_load_path(str(tmp_path / "missing"))


@pytest.mark.skipif(not os.path.isdir("/root/reference/humanoid"), reason="the reference is only present in the build container")
def test_get_load_path_equals_the_references(tmp_path):
    """The reference's own function (loaded from its source file without importing isaacgym) on the same tree."""
    import ast
    src = open("/root/reference/humanoid/utils/helpers.py", encoding="utf-8").read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "get_load_path")
    ns = {"os": os}
    exec(compile(ast.Module([fn], []), "reference_get_load_path", "exec"), ns)
    root = str(tmp_path)
    _tree(root)
    for kw in ({}, {"load_run": "2024-01-01_10-00-00ti5"}, {"checkpoint": 900}, {"load_run": "2024-01-01_10-00-00ti5", "checkpoint": 100}):
        assert get_load_path(root, **kw) == ns["get_load_path"](root, **kw), kw


def test_update_cfg_from_args_sets_the_resume_fields():
    from ti5_isaacgym_b200.envs import DHT1StandCfgPPO
    train = DHT1StandCfgPPO()            # an INSTANCE, like the registry holds: its nested configs are its own
    assert train.runner.resume is False and train.runner.load_run == -1 and train.runner.checkpoint == -1
    _, train = update_cfg_from_args(None, train, default_args(resume=True, load_run="runA", checkpoint=500, max_iterations=7))
    assert train.runner.resume is True and train.runner.load_run == "runA" and train.runner.checkpoint == 500
    assert train.runner.max_iterations == 7
    assert DHT1StandCfgPPO().runner.resume is False          # other instances are untouched


class _FakeRunner:
    made = []

    def __init__(self, env, cfg, log_dir, device="cpu"):
        self.env, self.cfg, self.log_dir, self.device, self.loaded = env, cfg, log_dir, device, None
        _FakeRunner.made.append(self)

    def load(self, path, load_optimizer=True):
        self.loaded = (path, load_optimizer)


def test_make_alg_runner_resumes_from_the_checkpoint(tmp_path, monkeypatch):
    """With `runner.resume` the registry looks the checkpoint up under `log_root` and calls
    `runner.load(path, load_optimizer=False)`; without it nothing is loaded."""
    from ti5_isaacgym_b200.envs import DHT1StandCfg, DHT1StandCfgPPO
    algo = types.ModuleType("humanoid.algo")
    algo.DHOnPolicyRunner = _FakeRunner
    pkg = types.ModuleType("humanoid")
    pkg.algo = algo
    monkeypatch.setitem(sys.modules, "humanoid", pkg)
    monkeypatch.setitem(sys.modules, "humanoid.algo", algo)
    root = str(tmp_path)
    _tree(root)
    reg = TaskRegistry()
    reg.env_cfg_for_wandb = DHT1StandCfg()
    env = types.SimpleNamespace(_materialize=True)
    runner, train_cfg, log_dir = reg.make_alg_runner(env, train_cfg=DHT1StandCfgPPO(), log_root=root,
                                                     storage="reference", args=default_args(rl_device="cpu"))
    assert runner.loaded is None and log_dir.startswith(root) and runner.cfg["runner_class_name"] == "DHOnPolicyRunner"
    runner, train_cfg, _ = reg.make_alg_runner(env, train_cfg=DHT1StandCfgPPO(), log_root=root, storage="reference",
                                               args=default_args(rl_device="cpu", resume=True, checkpoint=900))
    assert runner.loaded == (os.path.join(root, "2024-02-01_10-00-00ti5", "model_900.pt"), False)
    # the reference's storage copies the held observation late: an env that hands out ring views is refused
    with pytest.raises(ValueError, match="materialize_obs=True"):
        reg.make_alg_runner(types.SimpleNamespace(_materialize=False), train_cfg=DHT1StandCfgPPO(),
                            log_root=root, storage="reference", args=default_args(rl_device="cpu"))
