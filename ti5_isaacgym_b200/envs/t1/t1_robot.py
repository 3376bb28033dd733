"""Constants of the T1 robot model that the step math needs.

Isaac Gym derives these from `resources/robots/t1/urdf/t1.urdf` at asset-load time
(`legged_robot.py:1304-1324, 1404-1417`): with `collapse_fixed_joints=True` the 24 URDF
links collapse to 13 rigid bodies and the 12 revolute leg joints are the DOFs.  The URDF
and PhysX are out of scope, so the numbers are tabulated here.
"""
from types import SimpleNamespace

_SIDES = ("l", "r")
DOF_NAMES = tuple(f"leg_{s}{j}_joint" for s in _SIDES for j in range(1, 7))
BODY_NAMES = ("base_link",) + tuple(f"leg_{s}{j}_link" for s in _SIDES for j in range(1, 7))

# <limit lower upper effort velocity> of leg_l1..l6, leg_r1..r6
_LOWER = (-0.523, -0.174, -0.785, 0.0, -2.0, -3.0)
_UPPER = (0.523, 0.174, 0.785, 2.09, 2.0, 3.0)
DOF_LOWER = _LOWER + _LOWER
DOF_UPPER = _UPPER + _UPPER
DOF_EFFORT = (102.0, 102.0, 267.0, 267.0, 80.0, 40.0, 102.0, 102.0, 267.0, 267.0, 80.0, 40.2)
DOF_VELOCITY = (10.7, 11.7, 11.5, 11.5, 11.6, 9.85, 11.7, 11.7, 11.5, 11.5, 11.6, 9.85)


def body_indices(substring):
    return [i for i, n in enumerate(BODY_NAMES) if substring in n]


def robot_constants(cfg):
    """Indices and limits as the reference computes them from the asset + cfg."""
    a = cfg.asset
    pen, term = [], []
    for n in a.penalize_contacts_on:
        pen.extend(body_indices(n))
    for n in a.terminate_after_contacts_on:
        term.extend(body_indices(n))
    return SimpleNamespace(
        dof_names=DOF_NAMES, body_names=BODY_NAMES, num_dof=len(DOF_NAMES), num_bodies=len(BODY_NAMES),
        feet_indices=body_indices(a.foot_name), knee_indices=body_indices(getattr(a, "knee_name", "\0")),
        penalised_contact_indices=pen, termination_contact_indices=term,
        dof_lower=DOF_LOWER, dof_upper=DOF_UPPER, dof_effort=DOF_EFFORT, dof_velocity=DOF_VELOCITY)
