"""Fake `isaacgym.gymapi` — TEST INFRASTRUCTURE ONLY (see package docstring).

`FakeGym` hands out plain torch tensors with the shapes and AoS strides of the
Isaac Gym tensor API (root (N,13), dof (N*D,2), contact (N*NB,3),
rigid (N*NB,13)); every simulation call is a no-op, so the reference's step
math runs on whatever the test harness writes into those tensors.
"""
import os
import xml.etree.ElementTree as ET
from types import SimpleNamespace

import numpy as np
import torch

SIM_PHYSX = 1
SIM_FLEX = 0
ENV_SPACE = 0
KEY_ESCAPE = 0
KEY_V = 1


class Vec3:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def __iadd__(self, o):
        self.x += o.x
        self.y += o.y
        self.z += o.z
        return self


class Quat:
    def __init__(self, x=0.0, y=0.0, z=0.0, w=1.0):
        self.x, self.y, self.z, self.w = x, y, z, w


class Transform:
    def __init__(self, p=None, r=None):
        self.p = p or Vec3()
        self.r = r or Quat()


class _Bag(SimpleNamespace):
    pass


def PlaneParams():
    return _Bag(normal=Vec3(), static_friction=1.0, dynamic_friction=1.0, restitution=0.0)


def TriangleMeshParams():
    return _Bag(nb_vertices=0, nb_triangles=0, transform=Transform(),
                static_friction=1.0, dynamic_friction=1.0, restitution=0.0)


def HeightFieldParams():
    return _Bag(column_scale=1.0, row_scale=1.0, vertical_scale=1.0, nbRows=0, nbColumns=0,
                transform=Transform(), static_friction=1.0, dynamic_friction=1.0, restitution=0.0)


def AssetOptions():
    return _Bag()


def CameraProperties():
    return _Bag(width=0, height=0)


class SimParams:
    def __init__(self, dt=0.001, substeps=1, use_gpu_pipeline=False):
        self.dt = dt
        self.substeps = substeps
        self.use_gpu_pipeline = use_gpu_pipeline
        self.up_axis = 1
        self.gravity = Vec3(0.0, 0.0, -9.81)
        self.physx = _Bag(use_gpu=False, num_subscenes=0, num_threads=0)


class _Asset:
    """Revolute joints / collapsed bodies of a URDF (collapse_fixed_joints=True)."""

    def __init__(self, path):
        root = ET.parse(path).getroot()
        fixed_children = {j.find("child").get("link") for j in root.findall("joint")
                          if j.get("type") == "fixed"}
        self.body_names = [l.get("name") for l in root.findall("link")
                           if l.get("name") not in fixed_children]
        joints = [j for j in root.findall("joint") if j.get("type") != "fixed"]
        self.dof_names = [j.get("name") for j in joints]
        dt = np.dtype([(k, np.float32) for k in
                       ("lower", "upper", "velocity", "effort", "stiffness", "damping",
                        "friction", "armature")] + [("hasLimits", np.bool_), ("driveMode", np.int32)])
        self.dof_props = np.zeros(len(joints), dtype=dt)
        for i, j in enumerate(joints):
            lim = j.find("limit")
            for k in ("lower", "upper", "velocity", "effort"):
                self.dof_props[k][i] = float(lim.get(k))
            # joint friction / damping: the URDF's <dynamics> where it has one, else a nominal non-zero value, so that the
            # multipliers of lr:915-931 show in what the reference hands back to the simulator
            dyn = j.find("dynamics")
            self.dof_props["friction"][i] = float(dyn.get("friction", 0.0)) if dyn is not None and float(dyn.get("friction", 0.0)) else 0.05 + 0.01 * i
            self.dof_props["damping"][i] = float(dyn.get("damping", 0.0)) if dyn is not None and float(dyn.get("damping", 0.0)) else 0.5 + 0.1 * i


class FakeGym:
    """One fake sim per FakeGym instance; `acquire_gym()` returns a fresh one."""

    def __init__(self):
        self.asset = None
        self.num_envs = 0
        self.device = "cpu"
        self.tensors = {}
        self.calls = []          # log of indexed setters, for tests
        self.body_mass0 = 10.0   # base-link mass handed to _process_rigid_body_props

    # --- set-up ---------------------------------------------------------------------------
    def create_sim(self, compute_device, graphics_device, physics_engine, sim_params):
        self.device = "cuda:%d" % compute_device if sim_params.use_gpu_pipeline else "cpu"
        return self

    def add_ground(self, sim, params):
        pass

    def add_triangle_mesh(self, sim, vertices, triangles, params):
        pass

    def add_heightfield(self, sim, samples, params):
        pass

    def load_asset(self, sim, root, file, options):
        self.asset = _Asset(os.path.join(root, file))
        return self.asset

    def get_asset_dof_count(self, a):
        return len(a.dof_names)

    def get_asset_rigid_body_count(self, a):
        return len(a.body_names)

    def get_asset_dof_properties(self, a):
        return a.dof_props.copy()

    def get_asset_rigid_shape_properties(self, a):
        return [_Bag(friction=1.0, restitution=0.0)]

    def get_asset_rigid_body_names(self, a):
        return list(a.body_names)

    def get_asset_dof_names(self, a):
        return list(a.dof_names)

    def set_asset_rigid_shape_properties(self, a, props):
        pass

    def create_env(self, sim, lower, upper, per_row):
        self.num_envs += 1
        return self.num_envs - 1

    def create_actor(self, env, asset, pose, name, group, filt, seg=0):
        return 0

    def set_actor_dof_properties(self, env, actor, props):
        if self.tensors:          # after prepare_sim: a hot-path call (lr:939), not asset set-up
            self._log("set_actor_dof_properties", (int(env), props.copy()))

    def get_actor_dof_properties(self, env, actor):
        return self.asset.dof_props.copy()

    def get_actor_rigid_body_properties(self, env, actor):
        return [_Bag(mass=self.body_mass0 if i == 0 else 1.0, com=Vec3(),
                     inertia=_Bag(x=Vec3(1, 0, 0), y=Vec3(0, 1, 0), z=Vec3(0, 0, 1)))
                for i in range(len(self.asset.body_names))]

    def set_actor_rigid_body_properties(self, env, actor, props, recomputeInertia=False):
        pass

    def get_actor_rigid_shape_properties(self, env, actor):
        return [_Bag(friction=1.0, restitution=0.0)]

    def set_actor_rigid_shape_properties(self, env, actor, props):
        pass

    def find_actor_rigid_body_handle(self, env, actor, name):
        return self.asset.body_names.index(name)

    def prepare_sim(self, sim):
        n, d, b = self.num_envs, len(self.asset.dof_names), len(self.asset.body_names)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        self.tensors = dict(root=z(n, 13), dof=z(n * d, 2), contact=z(n * b, 3), rigid=z(n * b, 13))
        self.tensors["root"][:, 6] = 1.0
        self.tensors["rigid"][:, 6] = 1.0

    def create_camera_sensor(self, env, props):
        return 0

    def create_viewer(self, sim, props):
        return None

    # --- tensor API -------------------------------------------------------------------------
    def acquire_actor_root_state_tensor(self, sim):
        return self.tensors["root"]

    def acquire_dof_state_tensor(self, sim):
        return self.tensors["dof"]

    def acquire_net_contact_force_tensor(self, sim):
        return self.tensors["contact"]

    def acquire_rigid_body_state_tensor(self, sim):
        return self.tensors["rigid"]

    def _noop(self, *a, **k):
        return None

    fetch_results = _noop
    viewer_camera_look_at = _noop

    # every tensor-API call of the hot path is logged as (name, payload) — the lower boundary a drop-in env must
    # reproduce.  Payloads are cloned where a call hands data to the simulator.  `log_calls = False`
    # (the default) keeps only the indexed setters, which older tests read.
    log_calls = False

    def _log(self, name, payload=None):
        if self.log_calls:
            self.calls.append((name, payload))

    def refresh_dof_state_tensor(self, sim):
        self._log("refresh_dof_state_tensor")

    def refresh_actor_root_state_tensor(self, sim):
        self._log("refresh_actor_root_state_tensor")

    def refresh_net_contact_force_tensor(self, sim):
        self._log("refresh_net_contact_force_tensor")

    def refresh_rigid_body_state_tensor(self, sim):
        self._log("refresh_rigid_body_state_tensor")

    def set_dof_actuation_force_tensor(self, sim, torques):
        self._log("set_dof_actuation_force_tensor", torques.clone())

    def simulate(self, sim):
        self._log("simulate")

    def set_actor_root_state_tensor(self, sim, state):
        self._log("set_actor_root_state_tensor", state.clone())

    def apply_rigid_body_force_tensors(self, sim, forces, torques, space=ENV_SPACE):
        self._log("apply_rigid_body_force_tensors", (forces.clone(), torques.clone(), space))

    def set_dof_state_tensor_indexed(self, sim, state, ids, n):
        if self.log_calls:
            self.calls.append(("set_dof_state_tensor_indexed", (state.clone(), ids.clone(), int(n))))
        else:
            self.calls.append(("dof", ids.clone()))

    def set_actor_root_state_tensor_indexed(self, sim, state, ids, n):
        if self.log_calls:
            self.calls.append(("set_actor_root_state_tensor_indexed", (state.clone(), ids.clone(), int(n))))
        else:
            self.calls.append(("root", ids.clone()))


def acquire_gym():
    return FakeGym()
