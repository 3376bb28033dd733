/*
 * ti5_step.h — C ABI of libti5step.so: the per-step vectorised environment math of the
 * `t1_dh_stand` task (Robotics-Engineer-khy/ti5_isaacgym) as sm_100a CUDA kernels.
 *
 * The reference has no FFI: its "operator API" is the Python `LeggedRobot` / `T1DHStandEnv`
 * methods.  Each entry point below replaces the torch-op chain of the method(s) cited next to
 * it ("lr" = humanoid/envs/base/legged_robot.py, "t1" = humanoid/envs/t1/t1_dh_stand_env.py,
 * "rs" = humanoid/algo/ppo/rollout_storage.py).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned (torch) memory, borrowed for the
 *     call; the library allocates nothing and keeps no pointers;
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous on it;
 *   - return value 0 = OK, otherwise a negative TI5_E* code; ti5_last_error() describes it;
 *     nothing throws across the ABI; there is NO CPU fallback;
 *   - one policy step is the call sequence
 *         ti5_begin_step                                                (or ti5_first_substep = clip + torque 0)
 *         DEC x { ti5_torque_substep ; <simulate> ; ti5_lag_push }      (or the fused ti5_substep)
 *         [ti5_sample_heights] ; ti5_post_physics ; ti5_reset_observe ; [ti5_materialize_obs]
 *     and all per-step counters live in device memory (Ti5Globals), so the sequence can be
 *     captured once into a CUDA graph and replayed.
 */
#ifndef TI5_STEP_H
#define TI5_STEP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TI5_ABI_VERSION 10

#define TI5_NUM_DOF 12      /* leg_l1..l6, leg_r1..r6 */
#define TI5_NUM_BODIES 13   /* base_link + 12 leg links after collapse_fixed_joints */
#define TI5_NUM_TERMS 28    /* every _reward_* the task defines, alphabetical (see ti5_reward_name) */
#define TI5_MAX_GAITS 4
#define TI5_MAX_WINDOWS 8   /* entries of push_duration / add_duration */
#define TI5_LOG_ROWS 64     /* ring of per-step `extras["episode"]` snapshots */
#define TI5_LOG_COLS 32     /* TI5_NUM_TERMS means + terrain_level + max_command_x + n_reset + spare */

/* error codes */
#define TI5_OK 0
#define TI5_EINVAL (-1)
#define TI5_ECUDA (-2)

/* Ti5Params.flags: the cfg booleans the kernels branch on (t1_cfg domain_rand / terrain / noise) */
enum {
  TI5_F_ADD_LAG = 1 << 0,            /* domain_rand.add_lag */
  TI5_F_ADD_DOF_LAG = 1 << 1,        /* domain_rand.add_dof_lag */
  TI5_F_ADD_IMU_LAG = 1 << 2,        /* domain_rand.add_imu_lag */
  TI5_F_RAND_GAINS = 1 << 3,         /* randomize_gains */
  TI5_F_RAND_COULOMB = 1 << 4,       /* randomize_coulomb_friction */
  TI5_F_RAND_TORQUE = 1 << 5,        /* randomize_torque */
  TI5_F_RAND_MOTOR_OFFSET = 1 << 6,  /* randomize_motor_offset */
  TI5_F_RAND_ARMATURE = 1 << 7,      /* randomize_joint_armature(_each_joint) */
  TI5_F_ADD_NOISE = 1 << 8,          /* noise.add_noise */
  TI5_F_MEASURE_HEIGHTS = 1 << 9,    /* terrain.measure_heights */
  TI5_F_PUSH_ROBOTS = 1 << 10,       /* domain_rand.push_robots */
  TI5_F_ADD_EXT_FORCE = 1 << 11,     /* domain_rand.add_ext_force */
  TI5_F_ONLY_POSITIVE = 1 << 12,     /* rewards.only_positive_rewards */
  TI5_F_CUSTOM_ORIGINS = 1 << 13,    /* mesh_type in {heightfield, trimesh} */
  TI5_F_TERRAIN_CURRICULUM = 1 << 14,
  TI5_F_COMMAND_CURRICULUM = 1 << 15,
  TI5_F_TRIMESH = 1 << 16,           /* log terrain_level in extras */
  TI5_F_RAND_LAG_STEPS = 1 << 17,    /* randomize_lag_timesteps */
  TI5_F_RAND_DOF_LAG_STEPS = 1 << 18,
  TI5_F_RAND_IMU_LAG_STEPS = 1 << 19,
  TI5_F_PLANE = 1 << 20,             /* heights are identically zero on a plane (lr:1564) */
  TI5_F_HEADING_COMMAND = 1 << 21,   /* commands.heading_command: the schedule draws a heading target (column 3), the yaw
                                        rate (column 2) follows the heading error on every step (t1:141-176, 185-188) */
  TI5_F_NO_SW_SWITCH = 1 << 22,      /* commands.sw_switch == False: the gait phase follows episode_length_buf and standing
                                        envs keep cycling (t1:89-90) */
  /* options t1_cfg marks "always False" (t1_cfg:290-312) */
  TI5_F_LAG_PERSTEP = 1 << 23,       /* randomize_lag_timesteps_perstep: action lag re-drawn every substep (lr:1038-1043);
                                        unfused substep kernels only (ti5_fused_step refuses) */
  TI5_F_DOF_LAG_PERSTEP = 1 << 24,   /* randomize_dof_lag_timesteps_perstep (t1:408-413) */
  TI5_F_IMU_LAG_PERSTEP = 1 << 25,   /* randomize_imu_lag_timesteps_perstep (t1:437-442) */
  TI5_F_POS_VEL_LAG = 1 << 26,       /* add_dof_pos_vel_lag without add_dof_lag (t1:416-431): joint positions and velocities
                                        lagged separately; both are read from the DOF ring (set TI5_F_ADD_DOF_LAG with it:
                                        the ring is pushed, dof_lag_len covers both ranges) */
  TI5_F_RAND_POS_LAG_STEPS = 1 << 27, TI5_F_RAND_VEL_LAG_STEPS = 1 << 28,   /* randomize_dof_{pos,vel}_lag_timesteps */
  TI5_F_POS_LAG_PERSTEP = 1 << 29, TI5_F_VEL_LAG_PERSTEP = 1 << 30          /* ..._perstep (t1:417-430) */
};

/* gait kinds of cfg.commands.gait (t1:138-177) */
enum { TI5_GAIT_STAND = 0, TI5_GAIT_WALK_SAGITTAL = 1, TI5_GAIT_WALK_LATERAL = 2, TI5_GAIT_ROTATE = 3,
       TI5_GAIT_WALK_OMNI = 4 };

/* phases of ti5_substep / ti5_reset_observe, options of ti5_post_physics (bit masks).
 * *_CHAINED: launch with a programmatic dependency on the preceding kernel of the stream, which must be the previous
 * kernel of the fused step sequence (ti5_first_substep / ti5_substep / ti5_sample_heights / ti5_post_physics): the
 * kernel becomes resident early, does the part of its work that depends on no other kernel of the step, and waits
 * for the predecessor to complete before anything else.  Never chain across a simulator call. */
enum { TI5_SUB_PUSH = 1, TI5_SUB_TORQUE = 2, TI5_SUB_CHAINED = 4 };
enum { TI5_RO_RESET = 1, TI5_RO_OBSERVE = 2, TI5_RO_CHAINED = 4 };
enum { TI5_POST_PUSH_LAST = 1, TI5_POST_CHAINED = 2 };
/* ti5_fused_step: FUSED_CHAINED = the launch follows ti5_sample_heights of the same step on the stream */
enum { TI5_FUSED_CHAINED = 2 };

/* how `tensor / python_scalar` is rounded: torch-CPU divides, torch-CUDA multiplies by the
 * reciprocal (ATen div_true_kernel_cuda); the reference therefore differs by device. */
enum { TI5_DIV_IEEE = 0, TI5_DIV_RECIPROCAL = 1 };

/* randomness: uniforms supplied by the caller (parity mode) or Philox4x32-10 in-kernel */
enum { TI5_RNG_POOLS = 0, TI5_RNG_PHILOX = 1 };

/* ---- scalars derived once from the config (lr:94-113, 212-249, 352-384; t1:326-357) ---------- */
typedef struct Ti5Params {
  int32_t num_envs;
  int32_t frame_stack;     /* H: long observation history (66) */
  int32_t c_frame_stack;   /* CH: critic frame stack (3) */
  int32_t num_single_obs;  /* K = 47 */
  int32_t priv_frame;      /* P = 73, or 73 + num_height_points with measure_heights */
  int32_t decimation;      /* DEC = 10 */
  int32_t lag_len;         /* lag_timesteps_range[1] + 1 = 31 */
  int32_t dof_lag_len;     /* 31 */
  int32_t imu_lag_len;     /* 11 */
  int32_t flags;           /* TI5_F_* */
  int32_t div_mode;        /* TI5_DIV_* */
  int32_t rng_mode;        /* TI5_RNG_* */
  int32_t env_block;       /* threads (= envs) per CTA of the per-env kernels: 32, 64 or 128 */
  int32_t num_gaits;
  int32_t gait_kind[TI5_MAX_GAITS];
  int32_t num_height_points;
  int32_t height_rows, height_cols;
  int32_t feet[2], knees[2];
  int32_t term_body, pen_body;            /* termination / penalised contact body (base_link) */
  int32_t lag_range[3][2];                /* action / dof / imu lag index ranges */
  int32_t n_push_dur, n_add_dur;
  int32_t terrain_rows, terrain_cols, max_terrain_level;
  uint32_t term_mask;                     /* bit i set = reward term i has a non-zero scale */
  int32_t log_len;                        /* L: rows of the frame logs (0 = no log); must be >= rollout length + H */
  int32_t applied_stride;                 /* floats between the rows of consecutive envs in Ti5Buffers.applied_force /
                                             applied_torque: 3 for plain (N,3) arrays, 3 * TI5_NUM_BODIES when they
                                             point at the body-0 rows of the (N,13,3) tensors that
                                             apply_rigid_body_force_tensors takes (t1:234-247) */
  int64_t max_episode_length;             /* ceil(episode_length_s / dt) = 2400 */
  int64_t push_interval, ext_force_interval, push_update_step, add_update_step;
  uint64_t seed;                          /* Philox key */

  float dt;                 /* decimation * sim dt (Python double, rounded to fp32 where torch does) */
  float cycle_time, action_scale, clip_actions, clip_obs, stand_threshold;
  float max_episode_length_s;
  float default_dof_pos[TI5_NUM_DOF], p_gains[TI5_NUM_DOF], d_gains[TI5_NUM_DOF];
  float torque_limits[TI5_NUM_DOF], dof_vel_limits[TI5_NUM_DOF];
  float reward_scale[TI5_NUM_TERMS];      /* float32(scale * dt), alphabetical term order */
  float noise_vec[64];                    /* noise_scale_vec (K entries) */
  float noise_level;
  float obs_lin_vel, obs_ang_vel, obs_dof_pos, obs_dof_vel, obs_quat, obs_height;
  float cmd_scale[3];
  float heading_w, heading_lo;            /* command_ranges["heading"] as (hi - lo, lo) */
  float base_init_state[13];
  /* reward constants (t1_cfg:360-381) */
  float base_height_target, foot_min_dist, foot_max_dist, knee_min_dist, knee_max_dist;
  float target_joint_pos_scale, target_joint_pos_scale2 /* float32(2 * scale) */, target_feet_height, target_feet_height_max, tracking_sigma;
  float max_contact_force, soft_dof_vel_limit;
  /* uniform -> value maps, stored as (hi - lo, lo) computed in double then rounded (torch_rand_float) */
  float torque_multi_w, torque_multi_lo;
  float motor_offset_w, motor_offset_lo;
  float kp_mult_w, kp_mult_lo, kd_mult_w, kd_mult_lo;
  float coulomb_w, coulomb_lo, viscous_w, viscous_lo;
  float armature_w[TI5_NUM_DOF], armature_lo[TI5_NUM_DOF];
  float dof_reset_w, dof_reset_lo;                 /* U(-0.1, 0.1), lr:1084 */
  float root_xy_w, root_xy_lo;                     /* lr:1105-1108 */
  float gait_time_w[TI5_MAX_GAITS], gait_time_lo[TI5_MAX_GAITS];
  float push_vel_w, push_vel_lo, push_ang_w, push_ang_lo;
  float ext_f_w[3], ext_f_lo[3], ext_t_w, ext_t_lo;
  float ext_force_div, ext_torque_div;             /* ext_force_max_x + 0.1, ext_torque_max + 0.1 */
  float border_size, horizontal_scale, vertical_scale, terrain_env_length;
  double push_duration[TI5_MAX_WINDOWS], add_duration[TI5_MAX_WINDOWS];   /* duration / dt, in steps (double) */
  double cmd_curriculum_max;                       /* commands.max_curriculum */
  double tracking_lin_vel_scale;                   /* reward_scales["tracking_lin_vel"] (double) */
  /* (appended: the fields above keep the offsets the t1 configuration was tuned with) */
  int32_t lag_range_pv[2][2];                      /* dof-position / dof-velocity lag index ranges (TI5_F_POS_VEL_LAG) */
  int32_t flags2;                                  /* TI5_F2_* (the 31 bits of `flags` are taken) */
  float joint_friction_w, joint_friction_lo, joint_damping_w, joint_damping_lo;   /* lr:762-773 multiplier ranges */
  int32_t pad_[6];                                 /* sizeof == 1472 = 23 x 64: the structs that follow this one in a kernel's
                                                      parameter space start on a constant-cache line */
} Ti5Params;

/* Ti5Params.flags2 */
enum {
  TI5_F2_RAND_JOINT_FRICTION = 1 << 0,   /* randomize_joint_friction (one multiplier per env; lr:762-763, 921-925) */
  TI5_F2_RAND_JOINT_DAMPING = 1 << 1     /* randomize_joint_damping (lr:772-773, 926-930) */
};

/* ---- device-resident counters and curriculum state (single instance per env object) -------- */
typedef struct Ti5Globals {
  int64_t step_index;          /* policy steps COMPLETED (advanced by the observation kernel, the last of a step) */
  int64_t step_now;            /* index of the step in progress, published by ti5_post_physics / ti5_reset_bookkeeping */
  int64_t common_step_offset;  /* common_step_counter = step_index + common_step_offset */
  int32_t n_reset;             /* envs reset in the current step (lr:490) */
  int32_t n_listed[2];         /* entries of Ti5Buffers.reset_list (arrival order), double-buffered by the parity of the
                                  step in progress: a step fills [step & 1] and zeroes [(step + 1) & 1] for its successor */
  int32_t tickets[2];          /* [0]: CTAs of ti5_reset_observe that have read step_index (the last one advances it);
                                  [1]: last-CTA-done counter of the terrain-level mean */
  int32_t is_first_add_force[2]; /* lr:90, t1:205-215; double-buffered by step parity (read [step&1], write [(step+1)&1]) */
  double cmd_range[2][3][2];   /* [step parity][lin_vel_x, lin_vel_y, ang_vel_yaw][lo, hi]; the command curriculum
                                  (lr:1160-1169) writes the next step's copy */
} Ti5Globals;

/* ---- device buffers.  (N,k) means row-major per-env rows.  See DESIGN.md for the layout ---- */
typedef struct Ti5Buffers {
  Ti5Globals* globals;
  /* gym tensor API, AoS (lr:137-154) */
  float* root_states;      /* (N,13) */
  float* dof_state;        /* (N,12,2) */
  float* contact_forces;   /* (N,13,3) */
  float* rigid_state;      /* (N,13,13) */
  /* actuation */
  float* actions;          /* (N,12) clipped actions of this step */
  float* torques;          /* (N,12) torques of the last substep evaluated (lr:401) */
  float* torques_substeps; /* optional (DEC,N,12), or NULL: ti5_fused_step leaves the torques of EVERY substep here — what
                              set_dof_actuation_force_tensor is handed DEC times per step (lr:403) */
  float* torque_multi;     /* (N,12) */
  float* p_gains_r;        /* (N,12) randomized_p_gains */
  float* d_gains_r;        /* (N,12) */
  float* motor_offsets;    /* (N,12) */
  float* coulomb;          /* (N,12) randomized_joint_coulomb */
  float* viscous;          /* (N,12) randomized_joint_viscous */
  float* joint_armatures;  /* (N,12) */
  /* lag rings, slot-major: (len, N, width); slot of push j is j % len */
  float* act_ring;         /* (lag_len, N, 12) */
  float* dof_ring;         /* (dof_lag_len, N, 24) = cat(q, qd) */
  float* imu_ring;         /* (imu_lag_len, N, 6) = cat(base_ang_vel, euler) */
  int32_t* lag_timestep;   /* (N,3): action / dof / imu lag index */
  int64_t* ring_stamp;     /* (N): pushes older than this index read as zero (reset) */
  /* previous-step state */
  float* last_actions;     /* (N,12) */
  float* last_last_actions;
  float* last_dof_vel;     /* (N,12) */
  float* last_root_vel;    /* (N,6) */
  /* commands and gait schedule */
  float* commands;         /* (N,4) */
  int64_t* episode_length_buf;  /* (N) */
  int64_t* phase_length_buf;    /* (N) */
  int32_t* gait_time;      /* (N,num_gaits) */
  float* gait_start;       /* (N) */
  /* feet bookkeeping (t1:642-657, 793-814) */
  float* feet_air_time;    /* (N,2) */
  float* feet_height;      /* (N,2) */
  float* last_feet_z;      /* (N,2) */
  uint8_t* last_contacts;  /* (N,2) bool */
  uint8_t* contact_filt;   /* (N,2) bool */
  /* derived base state (lr:475-481) */
  float* base_quat;        /* (N,4) */
  float* base_lin_vel;     /* (N,3) */
  float* base_ang_vel;     /* (N,3) */
  float* projected_gravity;/* (N,3) */
  float* base_euler_xyz;   /* (N,3) */
  float* feet_euler_xyz;   /* (N,2,3) */
  float* ref_dof_pos;      /* (N,12) */
  float* ref_action;       /* (N,12) */
  /* disturbances */
  float* ext_forces;       /* (N,3) */
  float* ext_torques;      /* (N,3) */
  float* rand_push_force;  /* (N,3) */
  float* rand_push_torque; /* (N,3) */
  float* applied_force;    /* force handed to apply_rigid_body_force_tensors for body 0: row e at e * applied_stride */
  float* applied_torque;   /* likewise */
  float* env_frictions;    /* (N) */
  float* body_mass;        /* (N) */
  /* terrain */
  float* env_origins;      /* (N,3) */
  int64_t* terrain_levels; /* (N) */
  int64_t* terrain_types;  /* (N) */
  float* terrain_origins;  /* (terrain_rows, terrain_cols, 3) */
  int16_t* height_samples; /* (height_rows, height_cols) */
  float* height_points;    /* (num_height_points, 2) base-frame scan grid (lr:1535-1549) */
  float* measured_heights; /* (N, num_height_points) */
  /* outputs */
  float* rew_buf;          /* (N) */
  uint8_t* reset_buf;      /* (N) bool */
  uint8_t* time_out_buf;   /* (N) bool */
  uint8_t* time_outs_latched; /* (N) bool: extras["time_outs"] (only rewritten on steps with a reset) */
  float* episode_sums;     /* (TI5_NUM_TERMS, N) */
  float* reward_terms;     /* (TI5_NUM_TERMS, N) scaled per-term rewards of this step, or NULL */
  int32_t* reset_ids;      /* (N) ascending ids of the envs reset this step */
  int32_t* reset_list;     /* (N) the same ids in arrival order (work list for the history clear) */
  float* dof_props;        /* optional (N, TI5_NUM_DOF, 3), or NULL: row r = [friction multiplier, damping multiplier,
                              armature] per DOF of env reset_ids[r], written next to the id list by the reset scatter —
                              lr:915-939 `_refresh_actor_dof_props` as one dense tensor (see ti5_gather_dof_props) */
  int32_t* block_counts;   /* scratch: (ceil(N/32) + 1) */
  float* block_sums;       /* scratch: (ceil(N/32), TI5_LOG_COLS) */
  float* extras_log;       /* (TI5_LOG_ROWS, TI5_LOG_COLS) */
  /* observation histories: mirrored rings (N, 2H, K) and (N, 2CH, P) */
  float* obs_ring;
  float* priv_ring;
  float* obs_out;          /* optional contiguous (N, H*K) for ti5_materialize_obs */
  float* priv_out;         /* optional contiguous (N, CH*P) */
  /* frame logs for the rollout storage (Ti5Params.log_len = L > 0, else NULL): every frame the observation kernel
   * appends is also kept, un-cleared, in row (step - 1) % L, with the number of frames of that step's window that
   * are not zeroed by an earlier reset (t1:556-559) */
  float* frame_log;        /* (N, L, K) */
  float* priv_log;         /* (N, L, P) */
  int16_t* valid_log;      /* (L, N) */
  int32_t* hist_valid;     /* (N) frames appended since the env's histories were last cleared, capped at H */
  /* optional mirror of the per-step scalar outputs for a host-side caller: [rew f32 (4N bytes) | reset bool (N) |
   * time_outs bool (N)], written by ti5_reset_observe.  Meant to point at mapped pinned host memory (the stores go
   * straight over the bus, no copy-engine hop); the `actions_in` of ti5_first_substep may likewise be host-mapped. */
  uint8_t* host_out;
  uint64_t* debug_ts;      /* optional (2, CTAs, 8) globaltimer probes of the two per-env kernels (profiling aid), or NULL */
  /* the lag options t1_cfg marks "always False" (appended, see Ti5Params) */
  int32_t* lag_pv;         /* (N,2): dof-position / dof-velocity lag index (TI5_F_POS_VEL_LAG), else NULL */
  int32_t* last_lag;       /* (2,N,5): the `last_*_lag_timestep` of the per-step re-draws, double-buffered (by substep
                              parity for the action lag, by step parity for the others); NULL without a *_PERSTEP flag */
  float* joint_coeffs;     /* (N,2): joint friction / damping multiplier of the env (TI5_F2_RAND_JOINT_*), else NULL; they
                              go to the simulator in columns 0 / 1 of `dof_props` */
  uint64_t pad_[6];        /* sizeof == 704 = 11 x 64 (see Ti5Params) */
} Ti5Buffers;

/* ---- caller-supplied uniforms of one step (TI5_RNG_POOLS).  All fp32 U[0,1) unless noted --- */
typedef struct Ti5Rng {
  const float* torque;        /* (DEC, N, 12)  lr:1071 */
  const float* cmd;           /* (2, 3, N, 3)  t1:126-177: [0] callback pass, [1] pass inside reset_idx */
  const float* push;          /* (N,5)   t1:223-226 */
  const float* ext;           /* (N,6)   t1:237-241 */
  const float* dofs;          /* (N,12)  lr:1084 */
  const float* root_xy;       /* (N,2)   lr:1105-1108 */
  const float* dr;            /* (N,7,12) lr:735-783 */
  const float* gait_time;     /* (N,3)   t1:116 */
  const float* noise;         /* (N,K)   t1:472 */
  const int64_t* lag_idx;     /* (N,3)   lr:608-629 integers already in range */
  const int64_t* gait_start;  /* (N)     t1:523 integers in {0,1} */
  const int64_t* terrain_level; /* (N)   lr:1156 integers in [0, max_terrain_level) */
  const int64_t* lag_idx_pv;  /* (N,2)   lr:639, 646 position / velocity lag at a reset, integers in range */
  const float* dr_joint;      /* (N,2)   lr:763, 773 joint friction / damping multiplier draws */
  const int64_t* lag_step;    /* (DEC+4, N) per-step re-draws, integers in range: rows 0..DEC-1 the action lag of each
                                 substep (lr:1039), then DOF, IMU, position, velocity (t1:409, 438, 418, 426) */
} Ti5Rng;

/* ---- entry points ------------------------------------------------------------------------- */

int ti5_version(void);
const char* ti5_last_error(void);
/* sizeof(Ti5Params), sizeof(Ti5Buffers), sizeof(Ti5Rng), sizeof(Ti5Globals): lets a binding check its mirror */
int ti5_struct_sizes(int32_t out[4]);
const char* ti5_reward_name(int term);

/* lr:393-394  `self.actions = clip(actions, +-clip_actions)` */
int ti5_begin_step(const Ti5Params* p, const Ti5Buffers* b, const float* actions_in, void* stream);
/* the same clip fused with the torque of substep 0 (one launch less per step) */
int ti5_first_substep(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, const float* actions_in, void* stream);

/* lr:1019-1074 `_compute_torques` for substep `k` in [0, DEC) */
int ti5_torque_substep(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, int k, void* stream);
/* lr:412-434 DOF- and IMU-lag push after the simulator substep `k` */
int ti5_lag_push(const Ti5Params* p, const Ti5Buffers* b, int k, void* stream);
/* fused: [push of substep k-1] + [torque of substep k] in one launch (phases = TI5_SUB_*) */
int ti5_substep(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, int k, int phases, void* stream);

/* lr:1551-1587 `_get_heights` (+ utils/math.py:8-12 quat_apply_yaw) */
int ti5_sample_heights(const Ti5Params* p, const Ti5Buffers* b, void* stream);

/* lr:458-489 + t1:179-215 + lr:509-517 + lr:654-680 + t1:572-946: counters, derived base state,
 * command schedule, push / external-force windows, termination, the reward sum, and the
 * compaction bookkeeping (n_reset, per-CTA offsets, episode statistics, command curriculum).
 * `options` = TI5_POST_*: PUSH_LAST fuses the lag push of the last substep into the same launch. */
int ti5_post_physics(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, int options, void* stream);

/* One launch for everything of a policy step that precedes the resets, for the case that NO simulator runs between
 * the substeps (synthetic state, or a simulator stepped elsewhere): lr:393-434 — action clip, DEC x [_compute_torques,
 * DOF- and IMU-lag push] — and then ti5_post_physics, i.e. the call sequence
 *     ti5_first_substep ; (DEC-1) x ti5_substep(PUSH|TORQUE) ; ti5_post_physics(PUSH_LAST)
 * with bit-identical results (same Philox counters, same op order): every dependency among those launches is per env,
 * so one CTA carries its envs through all of them — the joint state, gains and offsets are read once, the lagged
 * action rows a substep pushed itself come from registers, every substep's torques are still stored in turn.
 * Follow with ti5_reset_observe(TI5_RO_RESET | TI5_RO_OBSERVE | TI5_RO_CHAINED): 2 launches per step instead of 12. */
int ti5_fused_step(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, const float* actions_in, int options,
                   void* stream);

/* The same bookkeeping for an explicit `reset_idx(env_ids)` (lr:450-455 `reset()`): the caller wrote the
 * mask into reset_buf; follow with ti5_reset_scatter. */
int ti5_reset_bookkeeping(const Ti5Params* p, const Ti5Buffers* b, void* stream);

/* lr:490 `reset_buf.nonzero()`: stand-alone ascending compaction of any (N) bool mask */
int ti5_compact_resets(const uint8_t* mask, int32_t n, int32_t* ids_out, int32_t* count_out,
                       int32_t* scratch /* ceil(n/1024)+1 ints */, void* stream);

/* t1:483-559 `reset_idx` for the flagged envs and t1:368-481 `compute_observations` + lr:496-499
 * `last_*` copies, in one launch (phases = TI5_RO_*).  Must follow ti5_post_physics. */
int ti5_reset_observe(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, int phases, void* stream);
int ti5_reset_scatter(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, void* stream);
int ti5_observations(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, void* stream);

/* lr:915-939 `_refresh_actor_dof_props(env_ids)`: the reference walks env_ids in a Python loop and reads every
 * randomised joint property of every env with a device->host sync per element.  This gathers them for the envs
 * ids[0 .. min(*count, capacity)) into ONE dense (capacity, TI5_NUM_DOF, 3) tensor [friction multiplier, damping
 * multiplier, armature] (lr:921-936; t1 randomises the armature only, the multipliers are 1), so that the simulator
 * side needs a single copy (or none, if it takes device tensors).  `ids` / `count` are typically Ti5Buffers.reset_ids
 * and &Ti5Globals.n_reset: no host round trip to learn how many envs were reset. */
int ti5_gather_dof_props(const Ti5Params* p, const Ti5Buffers* b, const int32_t* ids, const int32_t* count,
                         int32_t capacity, float* props_out, void* stream);

/* lr:441-446: copy the current history windows into contiguous (N,H*K) / (N,CH*P) tensors */
int ti5_materialize_obs(const Ti5Params* p, const Ti5Buffers* b, void* stream);

/* rs:97-119 `RolloutStorage.compute_returns`: reverse GAE scan over (T,N) + advantage
 * normalisation with the unbiased std.  `stats` is scratch for the Welford partials:
 * 3 doubles per CTA + 4 (count, mean, M2 of the whole batch, written by ti5_gae_scan and
 * read by ti5_gae_normalize, so a multi-GPU caller can all-reduce them in between). */
int ti5_gae_scan(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                 float* returns, float* advantages, int32_t T, int32_t N, float gamma, float lam,
                 double* stats, int32_t* ticket, void* stream);
int ti5_gae_normalize(float* advantages, int32_t T, int32_t N, const double* stats, void* stream);
int ti5_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
            float* returns, float* advantages, int32_t T, int32_t N, float gamma, float lam,
            double* stats, int32_t* ticket, void* stream);

/* ---- rollout storage: rs:59-74 `add_transitions`, rs:129-173 `mini_batch_generator`, dh_ppo.py:93-103
 * `process_env_step`, dh_on_policy_runner.py:149-168 episode bookkeeping ------------------------------
 * The reference keeps every step's (N, H*K) observation window: (T,N,3102) fp32 = 2.4 GB at 8192 envs.  Here the
 * storage keeps only the per-step scalars and reads the windows back out of the env's frame logs (one K-float
 * frame per env and step) when a mini-batch is drawn.  All (T,N,.) arrays are row-major, flat index t*N + e as in
 * `flatten(0, 1)` (rs:134-150). */
typedef struct Ti5Rollout {
  int32_t num_envs, num_steps;                 /* N, T */
  int32_t frame_stack, c_frame_stack;          /* H, CH */
  int32_t num_single_obs, priv_frame;          /* K, P */
  int32_t log_len, num_actions;                /* L, A */
  const float* frame_log;                      /* (N, L, K)  written by ti5_reset_observe */
  const float* priv_log;                       /* (N, L, P) */
  const int16_t* valid_log;                    /* (L, N) */
  int32_t* frame_row;                          /* (T) log row holding the newest frame of step t's observation */
  float* actions;                              /* (T,N,A) */
  float* mu;                                   /* (T,N,A) */
  float* sigma;                                /* (T,N,A) */
  float* rewards;                              /* (T,N) */
  uint8_t* dones;                              /* (T,N) */
  float* values;                               /* (T,N) */
  float* actions_log_prob;                     /* (T,N) */
  float* returns;                              /* (T,N) */
  float* advantages;                           /* (T,N) */
  /* runner bookkeeping (optional, NULL to skip): running episode return / length per env and the finished
   * episodes of the rollout in (step, ascending env id) order — what the runner appends to rewbuffer / lenbuffer */
  float* cur_reward_sum;                       /* (N) */
  float* cur_episode_length;                   /* (N) */
  float* finished_rew;                         /* (T*N) */
  float* finished_len;                         /* (T*N) */
  int32_t* n_finished;                         /* [2], double-buffered by step parity; [0] zeroed by the caller at clear() */
} Ti5Rollout;

/* the (N,.) tensors of one transition (rs:4-19) as the policy and the env produced them */
typedef struct Ti5Transition {
  const float* actions;                        /* (N,A) */
  const float* action_mean;                    /* (N,A) */
  const float* action_sigma;                   /* (N,A) */
  const float* values;                         /* (N) */
  const float* actions_log_prob;               /* (N) */
  const float* rewards;                        /* (N) env rewards of the step, not yet bootstrapped */
  const uint8_t* dones;                        /* (N) bool */
  const uint8_t* time_outs;                    /* (N) bool, or NULL (no 'time_outs' in infos) */
} Ti5Transition;

/* the gathered rows of one mini-batch (rs:152-173); any pointer may be NULL to skip that column */
typedef struct Ti5Batch {
  float* obs;                                  /* (B, H*K) */
  float* critic_obs;                           /* (B, CH*P) */
  float* actions;                              /* (B,A) */
  float* values;                               /* (B) */
  float* advantages;                           /* (B) */
  float* returns;                              /* (B) */
  float* actions_log_prob;                     /* (B) */
  float* mu;                                   /* (B,A) */
  float* sigma;                                /* (B,A) */
} Ti5Batch;

/* sizeof(Ti5Rollout), sizeof(Ti5Transition), sizeof(Ti5Batch) */
int ti5_rollout_struct_sizes(int32_t out[3]);

/* dh_ppo.py:93-103 + rs:59-74 (+ runner :149-168) in one launch: rewards += gamma * values * time_outs, the row
 * `step` of every per-step array, the frame-log row of the step's observation, episode return / length accounting */
int ti5_store_transition(const Ti5Rollout* ro, const Ti5Transition* tr, int32_t step, int32_t frame_row, float gamma,
                         void* stream);
/* rs:152-164: rows `idx[0..B)` (flat t*N + e) of every column; the observation windows are rebuilt from the logs.
 * `order` (optional permutation of [0,B), or NULL) only sets which batch row each warp produces first: sorting the
 * rows by (env, step) lets neighbouring warps share the frames of overlapping windows; the output is unchanged. */
int ti5_gather_minibatch(const Ti5Rollout* ro, const int64_t* idx, const int32_t* order, int32_t B, const Ti5Batch* out,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TI5_STEP_H */
