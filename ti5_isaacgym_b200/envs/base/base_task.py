"""`BaseTask`: device / buffer plumbing every task shares (reference: humanoid/envs/base/base_task.py).

The reference acquires Isaac Gym here (base_task.py:14) and renders; this build keeps the same
constructor signature and public buffers (base_task.py:55-74) but the simulator handle is any
object with the gym tensor API — `SyntheticGym` when Isaac Gym is not installed.  Rendering /
viewer code is out of scope.
"""
import torch

from ...sim.synthetic import SyntheticGym


def parse_device_str(device):
    s = str(device).lower()
    if ":" in s:
        kind, idx = s.split(":")
        return kind, int(idx)
    return s, 0


class BaseTask:
    def __init__(self, cfg, sim_params, physics_engine, sim_device, headless, gym=None):
        self.sim_params = sim_params
        self.physics_engine = physics_engine
        self.sim_device = sim_device
        kind, self.sim_device_id = parse_device_str(sim_device)
        self.headless = headless
        if kind != "cuda" or not getattr(sim_params, "use_gpu_pipeline", True):
            raise RuntimeError("ti5_isaacgym_b200 runs the step math in sm_100a CUDA kernels only: "
                               f"sim_device={sim_device!r} with use_gpu_pipeline=True is required (no CPU fallback)")
        self.device = f"cuda:{self.sim_device_id}"
        self.graphics_device_id = -1 if headless else self.sim_device_id
        self.num_envs = cfg.env.num_envs
        self.num_obs = cfg.env.num_observations
        self.num_short_obs = int(cfg.env.num_single_obs * cfg.env.short_frame_stack)
        self.num_privileged_obs = cfg.env.num_privileged_obs
        self.num_actions = cfg.env.num_actions
        self.num_single_obs = cfg.env.num_single_obs
        self.extras = {}
        self.gym = gym if gym is not None else SyntheticGym(self.num_envs, self.device)
        self.create_sim()
        self.enable_viewer_sync = True
        self.viewer = None

    def create_sim(self):
        raise NotImplementedError

    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def reset_idx(self, env_ids):
        raise NotImplementedError

    def reset(self):
        """lr:450-455: reset every robot, then one zero-action step."""
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        obs, privileged_obs, _, _, _ = self.step(torch.zeros(self.num_envs, self.num_actions, device=self.device))
        return obs, privileged_obs

    def step(self, actions):
        raise NotImplementedError

    def render(self, sync_frame_time=True):
        return None
