"""`task_registry` with the reference's surface (humanoid/utils/task_registry.py:16-148).

`make_env` builds the CUDA env; `make_alg_runner` instantiates the runner class named by the train
config from `humanoid.algo` when the reference's (unchanged) PPO code is importable — the dense
actor-critic / PPO update is out of scope and stays in PyTorch.
"""
import os
from datetime import datetime
from types import SimpleNamespace

from .. import LEGGED_GYM_ROOT_DIR
from ..sim.synthetic import SimParams
from .helpers import class_to_dict, get_load_path, set_seed, update_cfg_from_args


def default_args(**over):
    """The fields of `helpers.get_args()` the registry reads (isaacgym's CLI parser is not available)."""
    a = dict(task="t1_dh_stand", resume=False, experiment_name=None, run_name=None, load_run=None, checkpoint=None,
             headless=True, horovod=False, rl_device="cuda:0", sim_device="cuda:0", num_envs=None, seed=None,
             max_iterations=None, physics_engine=1, use_gpu_pipeline=True)
    a.update(over)
    return SimpleNamespace(**a)


class TaskRegistry:
    def __init__(self):
        self.task_classes, self.env_cfgs, self.train_cfgs = {}, {}, {}

    def register(self, name, task_class, env_cfg, train_cfg):
        self.task_classes[name], self.env_cfgs[name], self.train_cfgs[name] = task_class, env_cfg, train_cfg

    def get_task_class(self, name):
        return self.task_classes[name]

    def get_cfgs(self, name):
        env_cfg, train_cfg = self.env_cfgs[name], self.train_cfgs[name]
        env_cfg.seed = train_cfg.seed
        return env_cfg, train_cfg

    def make_env(self, name, args=None, env_cfg=None, **env_kwargs):
        args = args or default_args()
        if name not in self.task_classes:
            raise ValueError(f"Task with name: {name} was not registered")
        if env_cfg is None:
            env_cfg, _ = self.get_cfgs(name)
        env_cfg, _ = update_cfg_from_args(env_cfg, None, args)
        set_seed(env_cfg.seed)
        sim = class_to_dict(env_cfg.sim)
        sim_params = SimParams(dt=sim["dt"], use_gpu_pipeline=args.use_gpu_pipeline, substeps=sim.get("substeps", 1))
        task_class = self.get_task_class(name)
        if isinstance(task_class, str):            # lazily resolved so that registering needs no GPU
            from .. import envs
            task_class = getattr(envs, task_class)
        env = task_class(cfg=env_cfg, sim_params=sim_params, physics_engine=args.physics_engine,
                         sim_device=args.sim_device, headless=args.headless, **env_kwargs)
        self.env_cfg_for_wandb = env_cfg
        return env, env_cfg

    def make_alg_runner(self, env, name=None, args=None, train_cfg=None, log_root="default", storage="frame_log"):
        """`storage="frame_log"` (default) swaps `FrameLogRolloutStorage` into the runner's PPO object: the
        unmodified runner stores the observation it acted on one env step late (dh_ppo.py:88 -> rs:62), which a
        view into the env's history ring does not survive.  `storage="reference"` keeps the reference's own
        RolloutStorage and therefore needs an env built with `materialize_obs=True`."""
        args = args or default_args()
        if train_cfg is None:
            if name is None:
                raise ValueError("Either 'name' or 'train_cfg' must be not None")
            _, train_cfg = self.get_cfgs(name)
        _, train_cfg = update_cfg_from_args(None, train_cfg, args)
        stamp = datetime.now().strftime("%Y-%m-%d_%H-%M-%S")
        if log_root == "default":
            log_root = os.path.join(LEGGED_GYM_ROOT_DIR, "logs", train_cfg.runner.experiment_name, "exported_data")
        log_dir = None if log_root is None else os.path.join(log_root, stamp + train_cfg.runner.run_name)
        all_cfg = {**class_to_dict(train_cfg), **class_to_dict(self.env_cfg_for_wandb)}
        try:
            import humanoid.algo as algo          # the reference's unchanged PPO stack
        except ImportError as e:
            raise ImportError("make_alg_runner needs the reference's `humanoid.algo` package on sys.path: the "
                              "PPO runner / actor-critic are out of this build's scope") from e
        if storage not in ("frame_log", "reference"):
            raise ValueError("storage must be 'frame_log' or 'reference'")
        if storage == "reference" and not getattr(env, "_materialize", False):
            raise ValueError("the reference's RolloutStorage copies the observation one env step after the policy saw "
                             "it; build the env with materialize_obs=True, or use storage='frame_log'")
        runner = getattr(algo, all_cfg["runner_class_name"])(env, all_cfg, log_dir, device=args.rl_device)
        if storage == "frame_log":
            from ..algo.rollout_storage import install_frame_log_storage
            install_frame_log_storage(runner.alg, env)
            env._materialize = False       # this storage never reads the held observation: ring views are enough
        if train_cfg.runner.resume:        # task_registry.py:137-143
            resume_path = get_load_path(log_root, load_run=train_cfg.runner.load_run, checkpoint=train_cfg.runner.checkpoint)
            print(f"Loading model from: {resume_path}")
            runner.load(resume_path, load_optimizer=False)
        return runner, train_cfg, log_dir


task_registry = TaskRegistry()
