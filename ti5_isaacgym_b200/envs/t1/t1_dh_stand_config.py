"""`DHT1StandCfg` / `DHT1StandCfgPPO` — the `t1_dh_stand` task configuration.

Same attribute tree and values as `humanoid/envs/t1/t1_dh_stand_config.py` (only the
overrides of the base config are listed, exactly as the reference subclasses do);
`make_t1_cfg(frame_stack=H)` builds the history-length variants of BASELINE config 5.
"""
from ..base.base_config import build_cfg, ns, ns_new
from ..base.legged_robot_config import LeggedRobotCfg, LeggedRobotCfgPPO, _PHYSX, _lag

_LEG = ("l", "r")
_INIT_ANGLE = 0.3
_JOINT_ANGLES = {f"leg_{s}{j}_joint": a for s in _LEG
                 for j, a in zip(range(1, 7), (0, 0, -_INIT_ANGLE, _INIT_ANGLE * 2, -_INIT_ANGLE, 0))}
_KP = {f"{j}_joint": v * 1 for j, v in zip(range(1, 7), (50, 70, 90, 120, 50, 30))}
_KD = {f"{j}_joint": v for j, v in zip(range(1, 7), (5, 7, 9, 12, 5, 3))}

# reflected rotor inertia per leg joint, (nominal, lo factor, hi factor), t1_cfg:265-276
_ARMATURE = ((0.15, 0.8, 1.2), (0.15, 0.8, 1.2), (3.6, 0.5, 1.0), (3.6, 0.5, 1.0), (0.1, 0.5, 1.1), (0.028, 0.5, 1.5))
_ARMATURE_RANGES = {f"joint_{i + 1}_armature_range": [_ARMATURE[i % 6][0] * _ARMATURE[i % 6][1],
                                                      _ARMATURE[i % 6][0] * _ARMATURE[i % 6][2]]
                    for i in range(12)}

_TERRAIN_MIX = {"flat": 0.5, "rough flat": 0.3, "slope up": 0.1, "slope down": 0.1, "rough slope up": 0,
                "rough slope down": 0, "stairs up": 0, "stairs down": 0, "discrete": 0, "wave": 0}

_REWARD_SCALES = dict(
    joint_pos=4, feet_clearance=1, feet_contact_number=1.2, feet_air_time=1, foot_slip=-0.5,
    feet_distance=0.2, knee_distance=0.2, feet_rotation=0.8, feet_contact_forces=-0.01,
    tracking_lin_vel=1.5, tracking_ang_vel=0.8, vel_mismatch_exp=0.5, low_speed=0.2, track_vel_hard=0.5,
    default_joint_pos=1, orientation=1, base_height=0.2, base_acc=0.2,
    action_smoothness=-0.03, torques=-2e-7, dof_vel=-2e-5, dof_acc=-5e-7, collision=-1, stand_still=2.5)


def make_t1_cfg(frame_stack=66, c_frame_stack=3, name="DHT1StandCfg"):
    """Build the env config class.  `frame_stack` is the long-history length H: the
    observation is H x 47, the actor's Conv1d sees H input channels (t1_cfg:10-18)."""
    single_obs, single_priv = 47, 73
    return build_cfg(name, dict(
        env=ns(frame_stack=frame_stack, short_frame_stack=5, c_frame_stack=c_frame_stack, num_single_obs=single_obs,
               num_observations=int(frame_stack * single_obs), single_num_privileged_obs=single_priv,
               num_privileged_obs=int(c_frame_stack * single_priv), num_actions=12, num_envs=4096,
               episode_length_s=24, use_ref_actions=False, single_linvel_index=53, num_commands=5),
        safety=ns_new(pos_limit=1.0, vel_limit=1.0, torque_limit=0.85),
        asset=ns(file="{LEGGED_GYM_ROOT_DIR}/resources/robots/t1/urdf/t1.urdf", name="t1", foot_name="6_link",
                 knee_name="4_link", terminate_after_contacts_on=["base_link"], penalize_contacts_on=["base_link"],
                 self_collisions=0, flip_visual_attachments=False, replace_cylinder_with_capsule=False,
                 fix_base_link=False),
        terrain=ns(mesh_type="trimesh", curriculum=True, measure_heights=False, static_friction=0.6,
                   dynamic_friction=0.6, terrain_length=8, terrain_width=8, num_rows=20, num_cols=20,
                   max_init_terrain_level=5, platform=3, terrain_dict=dict(_TERRAIN_MIX),
                   terrain_proportions=list(_TERRAIN_MIX.values()), rough_flat_range=[0.005, 0.01],
                   slope_range=[0, 0.1], rough_slope_range=[0.005, 0.02], stair_width_range=[0.25, 0.25],
                   stair_height_range=[0.01, 0.1], discrete_height_range=[0.0, 0.01], restitution=0),
        noise=ns(add_noise=True, noise_level=1.5,
                 noise_scales=ns(dof_pos=0.02, dof_vel=1.5, ang_vel=0.2, lin_vel=0.1, quat=0.1, gravity=0.05,
                                 height_measurements=0.1)),
        init_state=ns(pos=[0.0, 0.0, 1.1], init_angle=_INIT_ANGLE, default_joint_angles=dict(_JOINT_ANGLES)),
        control=ns(control_type="P", stiffness=dict(_KP), damping=dict(_KD), action_scale=0.5, decimation=10),
        sim=ns(dt=0.001, substeps=1, up_axis=1, physx=ns(_PHYSX)),
        domain_rand=ns(
            randomize_friction=True, friction_range=[0.2, 1.3], restitution_range=[0.0, 0.4],
            push_robots=False, push_interval_s=6, update_step=2500 * 24,
            push_duration=[0, 0.05, 0.1, 0.15, 0.2, 0.25, 0.3], max_push_vel_xy=0.2, max_push_ang_vel=0.2,
            add_ext_force=True, ext_force_max_x=600, ext_force_max_y=400, ext_force_max_z=5, ext_torque_max=0,
            ext_force_interval_s=4, add_update_step=4000 * 24, add_duration=[0.0, 0.05, 0.1, 0.15],
            randomize_base_mass=True, added_mass_range=[-2.5, 2.5],
            randomize_com=True, com_displacement_range=[[-0.05, 0.05], [-0.05, 0.05], [-0.05, 0.05]],
            randomize_gains=True, stiffness_multiplier_range=[0.8, 1.2], damping_multiplier_range=[0.8, 1.2],
            randomize_torque=True, torque_multiplier_range=[0.8, 1.2],
            randomize_link_mass=True, added_link_mass_range=[0.9, 1.1],
            randomize_motor_offset=True, motor_offset_range=[-0.035, 0.035],
            randomize_joint_armature=True, randomize_joint_armature_each_joint=True,
            joint_armature_range=[0.001, 0.05], **_ARMATURE_RANGES,
            **_lag("lag", (0, 30), True), **_lag("dof_lag", (0, 30), True),
            add_dof_pos_vel_lag=False,
            **{k: v for k, v in _lag("dof_pos_lag", (7, 25)).items() if not k.startswith("add_")},
            **{k: v for k, v in _lag("dof_vel_lag", (7, 25)).items() if not k.startswith("add_")},
            **_lag("imu_lag", (0, 10), True),
            randomize_coulomb_friction=True, joint_coulomb_range=[0.1, 1.0], joint_viscous_range=[0.1, 0.9]),
        commands=ns(curriculum=True, max_curriculum=1.5, num_commands=4, resampling_time=25,
                    gait=["walk_omnidirectional", "stand", "walk_omnidirectional"],
                    gait_time_range={"walk_sagittal": [2, 6], "walk_lateral": [2, 6], "rotate": [2, 3],
                                     "stand": [2, 3], "walk_omnidirectional": [4, 6]},
                    stand_time=18, heading_command=False, stand_com_threshold=0.05, sw_switch=True,
                    ranges=ns_new(lin_vel_x=[-0.5, 0.5], lin_vel_y=[-0.5, 0.5], ang_vel_yaw=[-0.5, 0.5],
                              heading=[-3.14, 3.14])),
        rewards=ns_new(base_height_target=0.965, foot_min_dist=0.15, foot_max_dist=0.45, knee_min_dist=0.12,
                   knee_max_dist=0.35, target_joint_pos_scale=0.3, target_feet_height=0.02,
                   target_feet_height_max=0.08, cycle_time=0.8, only_positive_rewards=True, tracking_sigma=5,
                   max_contact_force=500, scales=ns_new(_REWARD_SCALES)),
        normalization=ns_new(obs_scales=ns_new(lin_vel=2, ang_vel=1, dof_pos=1, dof_vel=0.05, quat=1,
                                       height_measurements=5.0),
                         clip_observations=100, clip_actions=100),
    ), base=LeggedRobotCfg)


def make_t1_cfg_ppo(env_cfg_cls, name="DHT1StandCfgPPO"):
    """PPO/runner config; `lin_vel_idx` and `in_channels` are derived from the env config at
    class-build time like the reference does in its class bodies (t1_cfg:440, 460-465)."""
    e, t = env_cfg_cls.env, env_cfg_cls.terrain
    frame = e.single_num_privileged_obs + (t.num_height if t.measure_heights else 0)
    return build_cfg(name, dict(
        seed=5, runner_class_name="DHOnPolicyRunner",
        policy=ns(init_noise_std=1.0, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128],
                  state_estimator_hidden_dims=[256, 128, 64], kernel_size=[6, 4], filter_size=[32, 16],
                  stride_size=[3, 2], lh_output_dim=64, in_channels=e.frame_stack),
        algorithm=ns(entropy_coef=0.001, learning_rate=1e-5, num_learning_epochs=2, gamma=0.994, lam=0.9,
                     num_mini_batches=4, lin_vel_idx=frame * (e.c_frame_stack - 1) + e.single_linvel_index),
        runner=ns(policy_class_name="ActorCriticDH", algorithm_class_name="DHPPO", num_steps_per_env=24,
                  max_iterations=30000, save_interval=500, experiment_name="t1_dh_stand", run_name="ti5",
                  resume=False, load_run=-1, checkpoint=-1, resume_path=None),
    ), base=LeggedRobotCfgPPO)


DHT1StandCfg = make_t1_cfg()
DHT1StandCfgPPO = make_t1_cfg_ppo(DHT1StandCfg)
