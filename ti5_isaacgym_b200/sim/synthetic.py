"""Synthetic simulator state with the Isaac Gym tensor-API layout.

The only PhysX interaction of the hot path is the gym tensor API (lower
boundary): four AoS tensors that PhysX refreshes in place.  Where Isaac Gym is absent
(this repo's tests and benchmarks) the same tensors are filled with the seeded
near-nominal distribution of SURVEY.md 8(d), so the reference, the oracle and the CUDA
kernels all see identical inputs.

    root_states    (N, 13)      pos3 | quat_xyzw4 | lin_vel3 | ang_vel3     (lr:137)
    dof_state      (N*D, 2)     interleaved (pos, vel) per DOF              (lr:138-143)
    contact_forces (N*NB, 3)    net contact force per body                  (lr:151)
    rigid_state    (N*NB, 13)   pos3 | quat4 | lin_vel3 | ang_vel3 per body (lr:154)
"""
from types import SimpleNamespace

import torch

D, NB = 12, 13
FEET, KNEES = (6, 12), (4, 10)
_DEFAULT_POSE = (0.0, 0.0, -0.3, 0.6, -0.3, 0.0) * 2


def alloc_sim_tensors(num_envs, device):
    """Zeroed gym tensors (identity quaternions), as `acquire_*_tensor` would hand out."""
    z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)
    sim = SimpleNamespace(root_states=z(num_envs, 13), dof_state=z(num_envs * D, 2),
                          contact_forces=z(num_envs * NB, 3), rigid_state=z(num_envs * NB, 13))
    sim.root_states[:, 6] = 1.0
    sim.rigid_state[:, 6] = 1.0
    return sim


def _unit_quat(n, gen, device):
    q = torch.cat((0.03 * torch.randn(n, 3, generator=gen, device=device), torch.ones(n, 1, device=device)), 1)
    return q / q.norm(dim=1, keepdim=True)


def fill_synthetic_state(sim, env_origins, gen, base_contact_rate=0.01, device=None):
    """Overwrite the four gym tensors in place with one draw of the 8(d) distribution.
    `gen` is a torch.Generator living on the same device as the tensors it draws for
    (CPU generator + CPU tensors for the parity fixtures; CUDA for throughput runs)."""
    N = sim.root_states.shape[0]
    dev = sim.root_states.device if device is None else device
    rn = lambda *s: torch.randn(*s, generator=gen, device=dev)
    ru = lambda *s: torch.rand(*s, generator=gen, device=dev)
    tgt = sim.root_states.device
    default = torch.tensor(_DEFAULT_POSE, device=dev)

    dof = torch.stack((default + 0.05 * rn(N, D), 0.5 * rn(N, D)), dim=-1)
    root = torch.empty(N, 13, device=dev)
    root[:, 0:3] = env_origins.to(dev) + 0.1 * rn(N, 3)
    root[:, 2] = 0.96 + 0.02 * rn(N)
    root[:, 3:7] = _unit_quat(N, gen, dev)
    root[:, 7:13] = 0.3 * rn(N, 6)

    contact = torch.zeros(N, NB, 3, device=dev)
    in_contact = ru(N, 2) < 0.6
    f = torch.cat((20.0 * rn(N, 2, 2), (50.0 + 400.0 * ru(N, 2)).unsqueeze(-1)), dim=-1)
    f[..., 2] = torch.where(ru(N, 2) < 0.02, torch.full_like(f[..., 2], 700.0), f[..., 2])
    contact[:, list(FEET)] = f * in_contact.unsqueeze(-1)
    r = ru(N)
    contact[:, 0, 2] = torch.where(r < base_contact_rate, torch.full_like(r, 30.0), contact[:, 0, 2])
    contact[:, 0, 0] = torch.where((r >= base_contact_rate) & (r < base_contact_rate + 0.01),
                                   torch.full_like(r, 0.5), contact[:, 0, 0])

    rigid = torch.empty(N, NB, 13, device=dev)
    rigid[:, :, 0:3] = root[:, None, 0:3] + 0.05 * rn(N, NB, 3)
    for pair, half in ((FEET, 0.15), (KNEES, 0.12)):
        rigid[:, pair[0], 1] = root[:, 1] + half + 0.01 * rn(N)
        rigid[:, pair[1], 1] = root[:, 1] - half + 0.01 * rn(N)
    rigid[:, list(FEET), 2] = 0.05 + 0.04 * ru(N, 2) * (~in_contact)
    rigid[:, :, 3:7] = _unit_quat(N * NB, gen, dev).view(N, NB, 4)
    rigid[:, :, 7:13] = 0.3 * rn(N, NB, 6)

    sim.dof_state.copy_(dof.view(N * D, 2).to(tgt))
    sim.root_states.copy_(root.to(tgt))
    sim.contact_forces.copy_(contact.view(N * NB, 3).to(tgt))
    sim.rigid_state.copy_(rigid.view(N * NB, 13).to(tgt))
    return sim


def synthetic_actions(num_envs, gen, device):
    return 0.3 * torch.randn(num_envs, D, generator=gen, device=device)


def synthetic_height_field(rows=2100, cols=2100, seed=7, device="cpu"):
    """int16 height samples standing in for `Terrain.heightsamples` (lr:1237)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-20, 60, (rows, cols), generator=g, dtype=torch.int16).to(device)


class SimParams:
    """The two `gymapi.SimParams` fields the hot path reads (lr:96, base_task.py:26)."""

    def __init__(self, dt=0.001, use_gpu_pipeline=True, substeps=1):
        self.dt = dt
        self.use_gpu_pipeline = use_gpu_pipeline
        self.substeps = substeps


class SyntheticGym:
    """Stands where `gymapi.acquire_gym()` stands in the reference (base_task.py:14) when Isaac
    Gym is not installed: it owns the four state tensors with the tensor-API layout and accepts
    the tensor-API calls of the hot path (lr:403-410, 429, 464-466, 1088, 1118; t1:230, 247).
    `simulate` is a no-op unless a `physics` callable is installed (tests use it to refresh the
    state between substeps).

    Two capabilities a simulator binding can declare, both beyond Isaac Gym's own API (lower boundary):
      device_counts = True   the indexed setters take the id list as a fixed-capacity device buffer plus a DEVICE count
                             (`n` is a 0-dim int32 tensor), so a step needs no host round trip to learn `len(env_ids)`;
      set_actor_dof_properties_batched(sim, ids, props, n)   one (capacity, 12, 3) tensor [friction multiplier, damping
                             multiplier, armature] for the re-spawned envs instead of lr:915-939's per-env get/set loop.
    Without them the env falls back to what Isaac Gym offers: one 4-byte read-back of the count, then the calls of the
    reference with host integers."""

    device_counts = True
    needs_indexed_resets = False      # True: notify like Isaac Gym needs it (host count; the reference's exact call list)
    log_len = 64

    def __init__(self, num_envs, device):
        from collections import deque
        self.num_envs, self.device = num_envs, device
        self.tensors = alloc_sim_tensors(num_envs, device)
        self.physics = None
        self.substep = 0
        self.indexed_calls = deque(maxlen=self.log_len)      # (kind, int32 ids, n) of the indexed setters, newest last
        self.applied_forces = None
        self.pushed_root_states = None
        self.dof_props = None

    def acquire_actor_root_state_tensor(self, sim=None):
        return self.tensors.root_states

    def acquire_dof_state_tensor(self, sim=None):
        return self.tensors.dof_state

    def acquire_net_contact_force_tensor(self, sim=None):
        return self.tensors.contact_forces

    def acquire_rigid_body_state_tensor(self, sim=None):
        return self.tensors.rigid_state

    def simulate(self, sim=None):
        if self.physics is not None:
            self.physics(self.substep)
        self.substep += 1

    def _noop(self, *a, **k):
        return None

    fetch_results = refresh_dof_state_tensor = refresh_actor_root_state_tensor = _noop
    refresh_net_contact_force_tensor = refresh_rigid_body_state_tensor = _noop
    set_dof_actuation_force_tensor = _noop

    def set_actor_root_state_tensor(self, sim, state):
        self.pushed_root_states = state

    def set_dof_state_tensor_indexed(self, sim, state, ids, n):
        self.indexed_calls.append(("dof", ids, n))

    def set_actor_root_state_tensor_indexed(self, sim, state, ids, n):
        self.indexed_calls.append(("root", ids, n))

    def apply_rigid_body_force_tensors(self, sim, forces, torques, space=0):
        self.applied_forces = (forces, torques)

    def set_actor_dof_properties_batched(self, sim, ids, props, n):
        self.dof_props = (ids, props, n)


class RecordingGym(SyntheticGym):
    """A SyntheticGym that behaves like Isaac Gym's API surface (host counts, per-env DOF property structs, a physics
    callback between the substeps) and logs every tensor-API call as (name, payload): the lower boundary of one
    step, comparable call by call with what the reference issues against `oracle/shim` (tests)."""

    device_counts = False
    needs_indexed_resets = True
    set_actor_dof_properties_batched = None

    def __init__(self, num_envs, device):
        import numpy as np
        super().__init__(num_envs, device)
        self.calls = []
        self.physics = lambda k: None
        dt = np.dtype([(k, np.float32) for k in ("lower", "upper", "velocity", "effort", "stiffness", "damping",
                                                 "friction", "armature")])
        self._dof_props = [np.zeros(D, dtype=dt) for _ in range(num_envs)]

    def _log(self, name, payload=None):
        self.calls.append((name, payload))

    def simulate(self, sim=None):
        self._log("simulate")
        super().simulate(sim)

    def refresh_dof_state_tensor(self, sim=None):
        self._log("refresh_dof_state_tensor")

    def refresh_actor_root_state_tensor(self, sim=None):
        self._log("refresh_actor_root_state_tensor")

    def refresh_net_contact_force_tensor(self, sim=None):
        self._log("refresh_net_contact_force_tensor")

    def refresh_rigid_body_state_tensor(self, sim=None):
        self._log("refresh_rigid_body_state_tensor")

    def set_dof_actuation_force_tensor(self, sim, torques):
        self._log("set_dof_actuation_force_tensor", torques.clone())

    def set_actor_root_state_tensor(self, sim, state):
        self._log("set_actor_root_state_tensor", state.clone())

    def apply_rigid_body_force_tensors(self, sim, forces, torques, space=0):
        self._log("apply_rigid_body_force_tensors", (forces.clone(), torques.clone(), space))

    def set_dof_state_tensor_indexed(self, sim, state, ids, n):
        assert isinstance(n, int), "Isaac Gym takes the id count as a host integer"
        self._log("set_dof_state_tensor_indexed", (state.clone(), ids.clone(), n))

    def set_actor_root_state_tensor_indexed(self, sim, state, ids, n):
        assert isinstance(n, int)
        self._log("set_actor_root_state_tensor_indexed", (state.clone(), ids.clone(), n))

    def get_actor_dof_properties(self, env, actor):
        return self._dof_props[env].copy()

    def set_actor_dof_properties(self, env, actor, props):
        self._dof_props[env] = props.copy()
        self._log("set_actor_dof_properties", (int(env), props.copy()))


class SyntheticTerrain:
    """Stand-in for `humanoid.utils.terrain.Terrain` (one-time CPU set-up built on
    isaacgym.terrain_utils; out of scope): a random int16 height field of
    the real shape (2100 x 2100 for the t1 cfg) plus the (num_rows, num_cols, 3) platform-origin
    table the terrain curriculum indexes (lr:1158, 1493)."""

    def __init__(self, cfg, num_robots):
        import numpy as np
        self.cfg = cfg
        self.env_length, self.env_width = cfg.terrain_length, cfg.terrain_width
        px = int(cfg.terrain_length / cfg.horizontal_scale)
        py = int(cfg.terrain_width / cfg.horizontal_scale)
        border = int(cfg.border_size / cfg.horizontal_scale)
        self.tot_rows = int(cfg.num_rows * px) + 2 * border
        self.tot_cols = int(cfg.num_cols * py) + 2 * border
        self.heightsamples = synthetic_height_field(self.tot_rows, self.tot_cols, seed=7).numpy()
        self.vertices = np.zeros((3, 3), dtype=np.float32)
        self.triangles = np.zeros((1, 3), dtype=np.uint32)
        i, j = np.meshgrid(np.arange(cfg.num_rows), np.arange(cfg.num_cols), indexing="ij")
        self.env_origins = np.stack(((i + 0.5) * self.env_length, (j + 0.5) * self.env_width,
                                     0.01 * ((i * 7 + j * 3) % 11)), axis=-1).astype(np.float64)
