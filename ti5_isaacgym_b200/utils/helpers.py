"""Host helpers with the reference's names (`humanoid/utils/helpers.py`)."""
import os
import random

import numpy as np
import torch


def class_to_dict(obj):
    """Recursive attribute dump in `dir()` order (helpers.py:14-29).  The ALPHABETICAL key
    order is load-bearing: it fixes the order in which reward terms are evaluated and
    summed (legged_robot.py:357-378)."""
    if not hasattr(obj, "__dict__"):
        return obj
    out = {}
    for key in dir(obj):
        if key.startswith("_"):
            continue
        val = getattr(obj, key)
        out[key] = [class_to_dict(v) for v in val] if isinstance(val, list) else class_to_dict(val)
    return out


def update_class_from_dict(obj, d):
    for key, val in d.items():
        cur = getattr(obj, key, None)
        if isinstance(cur, type):
            update_class_from_dict(cur, val)
        else:
            setattr(obj, key, val)


def set_seed(seed):
    """helpers.py:42-64."""
    if seed == -1:
        seed = np.random.randint(0, 10000)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    return seed


def update_cfg_from_args(env_cfg, cfg_train, args):
    """helpers.py:125-148: only num_envs / seed / iteration / run-name overrides exist."""
    if env_cfg is not None and getattr(args, "num_envs", None) is not None:
        env_cfg.env.num_envs = args.num_envs
    if cfg_train is not None:
        if getattr(args, "seed", None) is not None:
            cfg_train.seed = args.seed
        for src, dst in (("max_iterations", "max_iterations"), ("experiment_name", "experiment_name"),
                         ("run_name", "run_name"), ("load_run", "load_run"), ("checkpoint", "checkpoint")):
            if getattr(args, src, None) is not None:
                setattr(cfg_train.runner, dst, getattr(args, src))
        if getattr(args, "resume", False):
            cfg_train.runner.resume = True
    return env_cfg, cfg_train


def get_load_path(root, load_run=-1, checkpoint=-1):
    """helpers.py:94-123: newest run directory under `root` (lexicographic order, "exported" skipped) unless `load_run`
    names one; in it the newest `model_<it>.pt` (zero-padded name order) unless `checkpoint` names the iteration."""
    try:
        runs = sorted(os.listdir(root))
        if "exported" in runs:
            runs.remove("exported")
        last_run = os.path.join(root, runs[-1])
    except (OSError, IndexError):
        raise ValueError("No runs in this directory: " + root)
    load_run = last_run if load_run == -1 else os.path.join(root, load_run)
    if checkpoint == -1:
        models = sorted((f for f in os.listdir(load_run) if "model" in f), key=lambda m: "{0:0>15}".format(m))
        model = models[-1]
    else:
        model = "model_{}.pt".format(checkpoint)
    return os.path.join(load_run, model)
