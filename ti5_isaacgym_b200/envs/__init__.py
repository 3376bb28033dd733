"""Task classes and configs, registered like `humanoid/envs/__init__.py:14` does."""
from .base.legged_robot_config import LeggedRobotCfg, LeggedRobotCfgPPO
from .t1.t1_dh_stand_config import DHT1StandCfg, DHT1StandCfgPPO, make_t1_cfg, make_t1_cfg_ppo


def __getattr__(name):
    # the env classes need the CUDA library; configs must stay importable without it
    if name in ("LeggedRobot", "T1DHStandEnv"):
        from .base.legged_robot import LeggedRobot
        from .t1.t1_dh_stand_env import T1DHStandEnv
        return {"LeggedRobot": LeggedRobot, "T1DHStandEnv": T1DHStandEnv}[name]
    raise AttributeError(name)


from ..utils.task_registry import task_registry  # noqa: E402

task_registry.register("t1_dh_stand", "T1DHStandEnv", DHT1StandCfg(), DHT1StandCfgPPO())
