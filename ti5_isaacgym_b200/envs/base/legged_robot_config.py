"""`LeggedRobotCfg` / `LeggedRobotCfgPPO` — defaults of the generic legged-robot task.

Attribute-for-attribute compatible with `humanoid/envs/base/legged_robot_config.py`
(the env, the registry and the PPO runner read these names); values are the reference's
defaults.  Written as a table and turned into the nested-class tree by `build_cfg`.
`tests/test_config.py` checks the tree against a committed dump of the reference's.
"""
from .base_config import BaseConfig, build_cfg, ns  # noqa: F401


def _linspace10(lo, n):
    """[-0.8, -0.7, ... ] in steps of 0.1 exactly as the literal decimal values."""
    return [round(lo + 0.1 * i, 1) + 0.0 for i in range(n)]


_PX = _linspace10(-0.8, 17)   # 1.6 m x 1.0 m height-scan grid, lr_cfg:29-30
_PY = _linspace10(-0.5, 11)

_TERRAIN_MIX = {"flat": 0.15, "rough flat": 0.15, "rough slope up": 0.0, "rough slope down": 0.0,
                "slope up": 0., "slope down": 0., "stairs up": 0.35, "stairs down": 0.25,
                "discrete": 0.0, "wave": 0.0}


def _per_joint(prefix, n, special, lo_hi, special_lo_hi):
    """joint_<i>_<prefix>_range entries: `special` joints get `special_lo_hi`."""
    return {f"joint_{i}_{prefix}_range": list(special_lo_hi if i in special else lo_hi)
            for i in range(1, n + 1)}


def _lag(prefix, rng, enabled=False):
    return {f"add_{prefix}": enabled, f"randomize_{prefix}_timesteps": True,
            f"randomize_{prefix}_timesteps_perstep": False, f"{prefix}_timesteps_range": list(rng)}


_DOMAIN_RAND = dict(
    randomize_friction=False, friction_range=[0.2, 1.3], restitution_range=[0.0, 0.4],
    push_robots=False, push_interval_s=4, update_step=2000 * 60, push_duration=[0, 0.1, 0.2, 0.3],
    max_push_vel_xy=0.2, max_push_ang_vel=0.2,
    add_ext_force=False, ext_force_max_xy=10, ext_force_max_z=5, ext_torque_max=0,
    ext_force_interval_s=10, add_update_step=2000 * 60, add_duration=[0, 0.1, 0.2, 0.3],
    continuous_push=False, max_push_force=0.5, max_push_torque=0.5,
    push_force_noise=0.5, push_torque_noise=0.5,
    randomize_base_mass=False, added_mass_range=[-2.5, 2.5],
    randomize_com=False, com_displacement_range=[[-0.05, 0.05]] * 3,
    randomize_link_com=False, link_com_displacement_range=[[-0.005, 0.005]] * 3,
    randomize_base_inertia=False, base_inertial_range=[[0.98, 1.02]] * 3,
    randomize_link_inertia=False, link_inertial_range=[[0.98, 1.02]] * 3,
    randomize_gains=False, stiffness_multiplier_range=[0.8, 1.2], damping_multiplier_range=[0.8, 1.2],
    randomize_torque=False, torque_multiplier_range=[0.8, 1.2],
    randomize_link_mass=False, added_link_mass_range=[0.9, 1.1],
    randomize_motor_offset=False, motor_offset_range=[-0.035, 0.035],
    randomize_joint_friction=False, randomize_joint_friction_each_joint=False,
    joint_friction_range=[0.01, 1.15],
    **_per_joint("friction", 10, (4, 5, 9, 10), (0.01, 1.15), (0.5, 1.3)),
    randomize_joint_damping=False, randomize_joint_damping_each_joint=False,
    joint_damping_range=[0.3, 1.5],
    **_per_joint("damping", 10, (4, 5, 9, 10), (0.3, 1.5), (0.9, 1.5)),
    randomize_joint_armature=False, randomize_joint_armature_each_joint=False,
    joint_armature_range=[0.0001, 0.05],
    **_per_joint("armature", 10, (), (0.0001, 0.05), ()),
    **_lag("lag", (5, 70)), **_lag("dof_lag", (0, 40)),
    add_dof_pos_vel_lag=False,
    **{k: v for k, v in _lag("dof_pos_lag", (7, 25)).items() if not k.startswith("add_")},
    **{k: v for k, v in _lag("dof_vel_lag", (7, 25)).items() if not k.startswith("add_")},
    **_lag("imu_lag", (1, 10)),
    randomize_coulomb_friction=False, joint_coulomb_range=[0.1, 0.9], joint_viscous_range=[0.10, 0.70],
)

_ZERO_SCALES = ("termination", "tracking_lin_vel", "tracking_ang_vel", "lin_vel_z", "ang_vel_xy", "orientation",
                "torques", "dof_vel", "dof_acc", "base_height", "feet_air_time", "collision", "feet_stumble",
                "action_rate", "stand_still")

_PHYSX = ns(num_threads=10, solver_type=1, num_position_iterations=4, num_velocity_iterations=0,
            contact_offset=0.01, rest_offset=0.0, bounce_threshold_velocity=0.5,
            max_depenetration_velocity=1.0, max_gpu_contact_pairs=2 ** 23,
            default_buffer_size_multiplier=5, contact_collection=2)

LeggedRobotCfg = build_cfg("LeggedRobotCfg", dict(
    env=ns(short_frame_stack=4, num_envs=4096, num_observations=235, num_privileged_obs=None, num_actions=12,
           env_spacing=3, send_timeouts=True, episode_length_s=20, num_commands=5, add_stand_bool=False,
           add_target_dof_scale=False),
    terrain=ns(mesh_type="trimesh", horizontal_scale=0.1, vertical_scale=0.005, border_size=25, curriculum=True,
               static_friction=1.0, dynamic_friction=1.0, restitution=0., measure_heights=False,
               measured_points_x=list(_PX), measured_points_y=list(_PY),
               measured_base_points_x=list(_PX), measured_base_points_y=list(_PY),
               measured_feet_points_x=list(_PX), measured_feet_points_y=list(_PY),
               num_height=len(_PX) * len(_PY), selected=False, terrain_kwargs=None, max_init_terrain_level=5,
               terrain_length=8., terrain_width=8., num_rows=10, num_cols=20, platform=3.,
               terrain_dict=dict(_TERRAIN_MIX), terrain_proportions=list(_TERRAIN_MIX.values()),
               rough_flat_range=[0.005, 0.02], slope_range=[0, 0.4], rough_slope_range=[0.005, 0.02],
               stair_width_range=[0.25, 0.25], stair_height_range=[0.04, 0.1],
               discrete_height_range=[0.05, 0.25], slope_treshold=0.75),
    commands=ns(curriculum=True, max_curriculum=1, num_commands=4, resampling_time=10, heading_command=True,
                ranges=ns(lin_vel_x=[-1.0, 1.0], lin_vel_y=[-1.0, 1.0], ang_vel_yaw=[-1, 1], heading=[-3.14, 3.14])),
    init_state=ns(pos=[0.0, 0.0, 1.0], rot=[0.0, 0.0, 0.0, 1.0], lin_vel=[0.0, 0.0, 0.0], ang_vel=[0.0, 0.0, 0.0],
                  default_joint_angles={"joint_a": 0, "joint_b": 0}),
    control=ns(control_type="P", stiffness={"joint_a": 10.0, "joint_b": 15.}, damping={"joint_a": 1.0, "joint_b": 1.5},
               action_scale=0.5, decimation=4),
    asset=ns(file="", name="legged_robot", foot_name="None", penalize_contacts_on=[], terminate_after_contacts_on=[],
             disable_gravity=False, collapse_fixed_joints=True, fix_base_link=False, default_dof_drive_mode=3,
             self_collisions=0, replace_cylinder_with_capsule=True, flip_visual_attachments=True, density=0.001,
             angular_damping=0, linear_damping=0, max_angular_velocity=1000, max_linear_velocity=1000, armature=0,
             thickness=0.01),
    domain_rand=ns(_DOMAIN_RAND),
    rewards=ns(scales=ns({k: (0.0 if k in ("tracking_lin_vel", "tracking_ang_vel", "feet_air_time") else -0.0)
                          for k in _ZERO_SCALES}),
               only_positive_rewards=True, tracking_sigma=0.25, max_contact_force=100.),
    normalization=ns(obs_scales=ns(lin_vel=2.0, ang_vel=0.25, dof_pos=1.0, dof_vel=0.05, height_measurements=5.0),
                     clip_observations=100., clip_actions=100.),
    noise=ns(add_noise=True, noise_level=1.0,
             noise_scales=ns(dof_pos=0.01, dof_vel=1.5, lin_vel=0.1, ang_vel=0.2, gravity=0.05,
                             height_measurements=0.1)),
    viewer=ns(ref_env=0, pos=[22, 3, 6], lookat=[0, 3, 0]),
    sim=ns(dt=0.005, substeps=1, gravity=[0., 0., -9.81], up_axis=1, physx=_PHYSX),
))

LeggedRobotCfgPPO = build_cfg("LeggedRobotCfgPPO", dict(
    seed=1, runner_class_name="OnPolicyRunner",
    policy=ns(init_noise_std=1.0, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128]),
    algorithm=ns(value_loss_coef=1.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.01,
                 num_learning_epochs=5, num_mini_batches=4, learning_rate=1.e-3, schedule="adaptive",
                 gamma=0.99, lam=0.95, desired_kl=0.01, max_grad_norm=1.),
    runner=ns(policy_class_name="ActorCritic", algorithm_class_name="PPO", num_steps_per_env=24,
              max_iterations=1500, save_interval=100, experiment_name="test", run_name="", resume=False,
              load_run=-1, checkpoint=-1, resume_path=None),
))
