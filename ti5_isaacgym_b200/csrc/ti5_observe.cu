// Post-physics phase, part 2: reset scatter for the envs flagged by ti5_post_physics, then the
// observation frames, the history rings and the `last_*` copies.  One thread per env; frames
// are staged in shared memory so the ring rows are written as contiguous 188 B / 292 B runs.
//
// Replaces t1:483-559 (T1 reset_idx with lr:604-651 randomize_lag_props, lr:732-783
// randomize_dof_props, lr:1076-1120 _reset_dofs/_reset_root_states, lr:1138-1158 terrain
// curriculum, t1:109-124 generate_gait_time), t1:368-481 (compute_observations with t1:250-274
// compute_ref_state), lr:496-499 (last_* copies) and lr:441-446 (observation clip).
//
// History layout: the reference keeps a deque of H (N,47) tensors and re-stacks it into a fresh
// (N, H*47) tensor every step (12.4 KB/env written, then clipped again).  Here the history is a
// mirrored ring (N, 2H, K): step s writes its frame at slots s%H and s%H+H, so the H most recent
// frames, oldest first, are always the contiguous run starting at slot s%H+1 of each env row —
// the caller returns that run as a strided view and nothing is re-stacked.  Frames are clipped to
// +-clip_obs on the way in: clip(stack(f)) == stack(clip(f)), and nothing else reads the ring.
#include "ti5_device.cuh"
#include "ti5_host.h"

namespace ti5 {

struct ObsRng {
  const Ti5Params& p;
  const Ti5Rng& r;
  uint64_t step;
  bool philox;
  // the eight per-DOF uniforms of a re-spawned env: [dof reset, torque multi, motor offset, kp, kd, coulomb,
  // viscous, armature] for DOF d (lr:1084, lr:735-783)
  __device__ __forceinline__ void dof_draws(int e, int d, float u[8], const float* pre = nullptr) const {
    if (pre) {                 // drawn before the grid wait (same Philox calls), parked in shared memory
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = pre[d * 8 + i];
    } else if (philox) {
      const float4 a = philox_u4(p.seed, step, S_DR, (e * D + d) * 2), c = philox_u4(p.seed, step, S_DR, (e * D + d) * 2 + 1);
      u[0] = a.x; u[1] = a.y; u[2] = a.z; u[3] = a.w; u[4] = c.x; u[5] = c.y; u[6] = c.z; u[7] = c.w;
    } else {
      u[0] = r.dofs[(size_t)e * D + d];
#pragma unroll
      for (int row = 0; row < 7; ++row) u[1 + row] = r.dr[((size_t)e * 7 + row) * D + d];
    }
  }
  __device__ __forceinline__ float root_xy(int e, int c) const {
    return philox ? philox_u(p.seed, step, S_ROOT, e * 2 + c) : r.root_xy[(size_t)e * 2 + c];
  }
  // the three command uniforms of gait `gi` (one Philox call)
  __device__ __forceinline__ void cmd3(int pass, int gi, int e, int N, float u[3]) const {
    if (philox) {
      const float4 a = philox_u4(p.seed, step, S_CMD + 4 * pass + gi, e);
      u[0] = a.x; u[1] = a.y; u[2] = a.z;
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) u[c] = r.cmd[((size_t)(pass * p.num_gaits + gi) * N + e) * 3 + c];
    }
  }
  // uniforms 4g .. 4g+3 of the env's noise row
  __device__ __forceinline__ float4 noise4(int e, int g4, int K) const {
    if (philox) {      // inlined: the caller unrolls four independent Philox chains side by side
      const uint4 w = philox4_inline(p.seed, step, S_NOISE, e * 16 + g4);
      return make_float4(u01(w.x), u01(w.y), u01(w.z), u01(w.w));
    }
    const float* row = r.noise + (size_t)e * K;
    const int k = 4 * g4;
    return make_float4(row[k], k + 1 < K ? row[k + 1] : 0.f, k + 2 < K ? row[k + 2] : 0.f, k + 3 < K ? row[k + 3] : 0.f);
  }
  // the schedule draws of a re-spawned env: three lag indices and the gait start (lr:608-629, t1:523) from
  // one Philox call, the gait durations (t1:116) from another
  __device__ __forceinline__ void schedule_draws(int e, int lag[3], float& gait_start, float gait_u[TI5_MAX_GAITS],
                                                 const float* pre = nullptr) const {
    if (philox) {
      float4 a, c;
      if (pre) { a = make_float4(pre[0], pre[1], pre[2], pre[3]); c = make_float4(pre[4], pre[5], pre[6], pre[7]); }
      else { a = philox_u4(p.seed, step, S_LAG, e); c = philox_u4(p.seed, step, S_GAIT_TIME, e); }
      const float ua[3] = {a.x, a.y, a.z};
#pragma unroll
      for (int w = 0; w < 3; ++w) {
        const int lo = p.lag_range[w][0], hi = p.lag_range[w][1];
        const int v = lo + (int)(ua[w] * (float)(hi - lo + 1));
        lag[w] = v > hi ? hi : v;
      }
      gait_start = a.w < 0.5f ? 0.0f : 0.5f;
      gait_u[0] = c.x; gait_u[1] = c.y; gait_u[2] = c.z; gait_u[3] = c.w;
    } else {
#pragma unroll
      for (int w = 0; w < 3; ++w) lag[w] = (int)r.lag_idx[(size_t)e * 3 + w];
      gait_start = (float)r.gait_start[e] * 0.5f;
      for (int gi = 0; gi < p.num_gaits; ++gi) gait_u[gi] = r.gait_time[(size_t)e * p.num_gaits + gi];
    }
  }
  // an integer in the lag range of `kind` (0 action, 1 DOF, 2 IMU, 3 position, 4 velocity) from one uniform, like the
  // reset draws above
  __device__ __forceinline__ int lag_from_u(int kind, float u) const {
    const int lo = kind < 3 ? p.lag_range[kind][0] : p.lag_range_pv[kind - 3][0];
    const int hi = kind < 3 ? p.lag_range[kind][1] : p.lag_range_pv[kind - 3][1];
    const int v = lo + (int)(u * (float)(hi - lo + 1));
    return v > hi ? hi : v;
  }
  // per-step re-draw of the DOF / IMU / position / velocity lag index (t1:409, 438, 418, 426), before the clamp
  __device__ __noinline__ int lag_step_draw(int kind, int e, int N) const {
    if (!philox) return (int)r.lag_step[(size_t)(p.decimation + kind - 1) * N + e];
    return lag_from_u(kind, philox_u(p.seed, step, S_LAGSTEP, e * 32 + 16 + kind));
  }
  // position / velocity lag index of a re-spawned env (lr:639, 646)
  __device__ __noinline__ int lag_pv_draw(int w, int e) const {
    if (!philox) return (int)r.lag_idx_pv[(size_t)e * 2 + w];
    return lag_from_u(3 + w, philox_u(p.seed, step, S_LAGSTEP, e * 32 + 24 + w));
  }
  // uniform of the env's joint friction (w = 0) / damping (w = 1) multiplier (lr:763, 773)
  __device__ __noinline__ float joint_coeff_u(int w, int e) const {
    return philox ? philox_u(p.seed, step, S_LAGSTEP, e * 32 + 26 + w) : r.dr_joint[(size_t)e * 2 + w];
  }
  __device__ __forceinline__ int64_t terrain_level(int e) const {
    if (!philox) return r.terrain_level[e] % p.max_terrain_level;
    int64_t v = (int64_t)(philox_u(p.seed, step, S_TERRAIN, e) * (float)p.max_terrain_level);
    return v >= p.max_terrain_level ? p.max_terrain_level - 1 : v;
  }
};

// t1:483-559 reset of env `es`, the part that is parallel over DOFs / reward terms: lanes 0-11 take one DOF
// each (joint state, the seven actuator draws, zeroed action history), lanes 0-27 clear one episode sum each.
// Runs once per flagged env of the warp, all 32 lanes together.
__device__ __forceinline__ void reset_env_dofs(const Ti5Params& p, const Ti5Buffers& b, const ObsRng& rng, int es, int lane,
                                               const float* pre) {
  const int N = p.num_envs;
  if (lane < D) {
    const int d = lane;
    const size_t o = (size_t)es * D + d;
    float u[8];
    rng.dof_draws(es, d, u, pre);
    // lr:1076-1090 joint state
    reinterpret_cast<float2*>(b.dof_state)[o] = make_float2(p.default_dof_pos[d] + affine(p.dof_reset_w, p.dof_reset_lo, u[0]), 0.0f);
    // lr:732-783 actuator randomisation
    if (p.flags & TI5_F_RAND_TORQUE) b.torque_multi[o] = affine(p.torque_multi_w, p.torque_multi_lo, u[1]);
    if (p.flags & TI5_F_RAND_MOTOR_OFFSET) b.motor_offsets[o] = affine(p.motor_offset_w, p.motor_offset_lo, u[2]);
    if (p.flags & TI5_F_RAND_GAINS) {
      b.p_gains_r[o] = affine(p.kp_mult_w, p.kp_mult_lo, u[3]) * p.p_gains[d];
      b.d_gains_r[o] = affine(p.kd_mult_w, p.kd_mult_lo, u[4]) * p.d_gains[d];
    }
    if (p.flags & TI5_F_RAND_COULOMB) {
      b.coulomb[o] = affine(p.coulomb_w, p.coulomb_lo, u[5]);
      b.viscous[o] = affine(p.viscous_w, p.viscous_lo, u[6]);
    }
    if (p.flags & TI5_F_RAND_ARMATURE) b.joint_armatures[o] = affine(p.armature_w[d], p.armature_lo[d], u[7]);
    // t1:513-518
    b.actions[o] = 0.0f; b.last_actions[o] = 0.0f; b.last_last_actions[o] = 0.0f; b.last_dof_vel[o] = 0.0f;
  }
  if (lane < TI5_NUM_TERMS && (p.term_mask & (1u << lane))) b.episode_sums[(size_t)lane * N + es] = 0.0f;   // t1:533
}

// Derived state of a re-spawned base (t1:548-552): base_init_state is the same for every env, so the rotations and
// Euler angles of the spawn pose are computed once per CTA, before the grid wait.
// Layout: quat(4) | base_lin_vel(3) | base_ang_vel(3) | projected_gravity(3) | euler(3)
constexpr int SPAWN_FLOATS = 16;
__device__ __forceinline__ void spawn_state(const Ti5Params& p, float* out) {
  const float* r0 = p.base_init_state;
  const float bq[4] = {r0[3], r0[4], r0[5], r0[6]};
  const V3 l = quat_rotate_inverse(bq, V3{r0[7], r0[8], r0[9]});
  const V3 a = quat_rotate_inverse(bq, V3{r0[10], r0[11], r0[12]});
  const V3 gr = quat_rotate_inverse(bq, V3{0.0f, 0.0f, -1.0f});
  float eul[3];
  euler_xyz(bq, eul);
  out[0] = bq[0]; out[1] = bq[1]; out[2] = bq[2]; out[3] = bq[3];
  out[4] = l.x; out[5] = l.y; out[6] = l.z;
  out[7] = a.x; out[8] = a.y; out[9] = a.z;
  out[10] = gr.x; out[11] = gr.y; out[12] = gr.z;
  out[13] = eul[0]; out[14] = eul[1]; out[15] = eul[2];
}

// The per-env scalar parts of the reset run on the env's own lane, so all flagged envs of a warp go in parallel.
// Base: terrain curriculum, root state, derived base quantities (`spawn`, above; `org` = the env's origin, loaded
// before the grid wait).
__device__ __forceinline__ void reset_env_base(const Ti5Params& p, const Ti5Buffers& b, const ObsRng& rng, int es,
                                               const float* spawn, float org[3]) {
  float* root = b.root_states + (size_t)es * RB;
  if (p.flags & TI5_F_TERRAIN_CURRICULUM) {                        // lr:1138-1158
    const float dx = root[0] - org[0], dy = root[1] - org[1];
    const float dist = sqrtf(dx * dx + dy * dy);
    const bool up = dist > (float)(p.terrain_env_length / 2.0);
    const float cx = b.commands[es * 4 + 0], cy = b.commands[es * 4 + 1];
    const bool down = (dist < sqrtf(cx * cx + cy * cy) * p.max_episode_length_s * 0.5f) && !up;
    int64_t lvl = b.terrain_levels[es] + (up ? 1 : 0) - (down ? 1 : 0);
    lvl = lvl >= p.max_terrain_level ? rng.terrain_level(es) : (lvl < 0 ? 0 : lvl);
    b.terrain_levels[es] = lvl;
    const float* o = b.terrain_origins + ((size_t)lvl * p.terrain_cols + b.terrain_types[es]) * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) { org[i] = o[i]; b.env_origins[es * 3 + i] = org[i]; }
  }
  float r0[RB];                                                     // lr:1092-1120
#pragma unroll
  for (int i = 0; i < RB; ++i) r0[i] = p.base_init_state[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) r0[i] += org[i];
  if (p.flags & TI5_F_CUSTOM_ORIGINS) {
    r0[0] += affine(p.root_xy_w, p.root_xy_lo, rng.root_xy(es, 0));
    r0[1] += affine(p.root_xy_w, p.root_xy_lo, rng.root_xy(es, 1));
  }
#pragma unroll
  for (int i = 0; i < RB; ++i) root[i] = r0[i];
  // t1:548-552 derived state of the re-spawned base
  reinterpret_cast<float4*>(b.base_quat)[es] = make_float4(spawn[0], spawn[1], spawn[2], spawn[3]);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    b.base_lin_vel[es * 3 + i] = spawn[4 + i];
    b.base_ang_vel[es * 3 + i] = spawn[7 + i];
    b.projected_gravity[es * 3 + i] = spawn[10 + i];
    b.base_euler_xyz[es * 3 + i] = spawn[13 + i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) b.last_root_vel[es * 6 + i] = 0.0f;
}

// Schedule (role 1): lag indices, counters, gait start and gait times.
template <bool LAGOPT>
__device__ __forceinline__ void reset_env_schedule(const Ti5Params& p, const Ti5Buffers& b, const ObsRng& rng, int es,
                                                   int64_t pushes, const float* pre) {
  int lag[3];
  float gs, gu[TI5_MAX_GAITS];
  rng.schedule_draws(es, lag, gs, gu, pre);
  // lr:604-633: the env's lag rings read as zero from now on; new lag indices
  b.ring_stamp[es] = pushes;
  if (p.flags & TI5_F_ADD_LAG) b.lag_timestep[es * 3 + 0] = (p.flags & TI5_F_RAND_LAG_STEPS) ? lag[0] : p.lag_range[0][1];
  if (p.flags & TI5_F_ADD_DOF_LAG) b.lag_timestep[es * 3 + 1] = (p.flags & TI5_F_RAND_DOF_LAG_STEPS) ? lag[1] : p.lag_range[1][1];
  if (p.flags & TI5_F_ADD_IMU_LAG) b.lag_timestep[es * 3 + 2] = (p.flags & TI5_F_RAND_IMU_LAG_STEPS) ? lag[2] : p.lag_range[2][1];
  if (LAGOPT && p.flags2) {                                                      // lr:762-763, 772-773: one multiplier per env
    if (p.flags2 & TI5_F2_RAND_JOINT_FRICTION)
      b.joint_coeffs[es * 2 + 0] = affine(p.joint_friction_w, p.joint_friction_lo, rng.joint_coeff_u(0, es));
    if (p.flags2 & TI5_F2_RAND_JOINT_DAMPING)
      b.joint_coeffs[es * 2 + 1] = affine(p.joint_damping_w, p.joint_damping_lo, rng.joint_coeff_u(1, es));
  }
  if (LAGOPT && (p.flags & TI5_F_LAG_OPTIONS)) {                                 // lr:610-611, 620-621, 630-650
    if (p.flags & TI5_F_POS_VEL_LAG) {
      b.lag_pv[es * 2 + 0] = (p.flags & TI5_F_RAND_POS_LAG_STEPS) ? rng.lag_pv_draw(0, es) : p.lag_range_pv[0][1];
      b.lag_pv[es * 2 + 1] = (p.flags & TI5_F_RAND_VEL_LAG_STEPS) ? rng.lag_pv_draw(1, es) : p.lag_range_pv[1][1];
    }
    const int perstep[5] = {TI5_F_LAG_PERSTEP, TI5_F_DOF_LAG_PERSTEP, TI5_F_IMU_LAG_PERSTEP, TI5_F_POS_LAG_PERSTEP,
                            TI5_F_VEL_LAG_PERSTEP};
    for (int k = 0; k < 5; ++k)
      if (p.flags & perstep[k]) {       // `last_*_lag_timestep[env_ids]` = the range maximum, in both copies
        const int hi = k < 3 ? p.lag_range[k][1] : p.lag_range_pv[k - 3][1];
        b.last_lag[((size_t)0 * p.num_envs + es) * 5 + k] = hi;
        b.last_lag[((size_t)1 * p.num_envs + es) * 5 + k] = hi;
      }
  }
  b.feet_air_time[es * 2 + 0] = 0.0f; b.feet_air_time[es * 2 + 1] = 0.0f;        // t1:519-523
  b.episode_length_buf[es] = 0;
  b.phase_length_buf[es] = 0;
  b.gait_start[es] = gs;
  // t1:109-124 generate_gait_time
  float rg[TI5_MAX_GAITS], sum = 0.0f;
  for (int gi = 0; gi < p.num_gaits; ++gi) {
    rg[gi] = affine(p.gait_time_w[gi], p.gait_time_lo[gi], gu[gi]);
    sum += rg[gi];
  }
  const float fac = (1.0f / sum) * (float)p.max_episode_length;   // Tensor.__rtruediv__: reciprocal, then multiply
  float run = (rg[0] * fac) * 0.0f;
  b.gait_time[es * p.num_gaits + 0] = (int32_t)run;
  for (int gi = 1; gi < p.num_gaits; ++gi) {
    run += rg[gi - 1] * fac;
    b.gait_time[es * p.num_gaits + gi] = (int32_t)run;
  }
}

// The CTA owns TB = env_block consecutive envs and runs 2, 4 or 8 "roles" x TB threads.  Roles 0 and 1 run the reset
// scatter (DOF-parallel part / per-env scalar parts); the observation frame (lagged proprioception, noise) and the
// privileged frame are built in eight parts dealt over the roles the CTA has, and appended to both rings by all of them
// — a frame set is 240 scalar stores per warp, the longest single-warp stretch of the kernel with two roles.
constexpr int OBS_ROLES = 2;
constexpr int DRAW_STRIDE = D * 8 + 8 + 1;     // floats per env of the parked reset draws (odd: conflict-free by env)

// MINB = 2 caps the kernel at 128 registers (122 used, no spills) for large grids, where four CTAs of 128 threads per SM
// instead of three are worth 10 us per step at 65536 envs; small grids keep the uncapped build (140 registers), which
// measured 1 us faster there.
// LAGOPT: the lag options t1_cfg marks "always False" (per-step re-draws, separate position / velocity lags) behind a
// compile-time switch — as run-time branches they cost the t1 configuration 1.2 us per step at 8192 envs (registers and
// stack of code that never runs).
template <int KC, int PC, int MINB, bool LAGOPT = false>
__global__ void __launch_bounds__(256, MINB)
reset_observe_kernel(const __grid_constant__ Ti5Params p, const __grid_constant__ Ti5Buffers b,
                     const __grid_constant__ Ti5Rng r, int phases) {
  extern __shared__ float smem[];      // per 32 envs: 32 x K observation frames, then 32 x P privileged frames
  __shared__ int s_warp[32];
  __shared__ float s_lvl[8];
  __shared__ bool s_last;

  // frame widths as compile-time constants where known: the flat ring-write loops divide by them
  const int N = p.num_envs, K = KC ? KC : p.num_single_obs, P = PC ? PC : p.priv_frame, H = p.frame_stack, CH = p.c_frame_stack;
  const int TB = p.env_block, tid = threadIdx.x;
  const int role = tid / TB, le = tid - role * TB;
  const int e = blockIdx.x * TB + le;
  const int lane = tid & 31, warp = tid >> 5, tile_warp = le >> 5;
  const int env_blocks = (N + TB - 1) / TB;              // CTAs beyond these only help clearing histories
  const bool live = e < N && blockIdx.x < env_blocks && role < OBS_ROLES;
  Ti5Globals* g = b.globals;
  // index of the step in progress: published by ti5_post_physics / ti5_reset_bookkeeping; a chained launch may not
  // read what its predecessor writes yet and is always part of a full step.
  // THIS kernel advances step_index, and a grid larger than one wave has CTAs that start after others have finished:
  // one thread per CTA reads the counter; behind the grid wait (the predecessor, which reads the counter too, is
  // complete there) it takes a ticket, and the CTA that draws the last one — every CTA of the grid has read by then —
  // publishes the new count.
  __shared__ int64_t s_step;
  __shared__ bool s_new_step;
  if (tid == 0) {
    const int64_t done = g->step_index;
    s_step = (phases & TI5_RO_CHAINED) ? done + 1 : g->step_now;
    s_new_step = s_step != done;       // false: compute_observations() called again on a completed step (same slots)
  }
  __syncthreads();
  const int64_t step = s_step;
  const int64_t pushes = step * p.decimation;            // lag pushes completed after this step
  const bool do_reset = (phases & TI5_RO_RESET) != 0, do_obs = (phases & TI5_RO_OBSERVE) != 0;
  const int dm = p.div_mode;

  // ---- reset bookkeeping left by ti5_post_physics: the per-CTA counts of flagged envs.  Every CTA derives
  // what it needs from them — the total (lr:490 `len(env_ids)`), its own offset into the ascending id list, and,
  // on the steps where it is due, the command-curriculum decision (lr:1160-1169) — so no CTA waits on another.
  __shared__ int s_tot[16], s_bef[16];
  __shared__ double s_trk[16];
  __shared__ double s_range[3][2];
  const ObsRng rng{p, r, (uint64_t)step, p.rng_mode == TI5_RNG_PHILOX};
  // staged frames: odd row strides (conflict-free).  With measured heights the privileged frame's 187 height entries are
  // NOT staged per env: the tile's rows of `measured_heights` are one contiguous block and arrive by a single bulk copy
  // into `s_h`, behind the frames of all the tile's envs
  const int npts_h = (p.flags & TI5_F_MEASURE_HEIGHTS) ? p.num_height_points : 0, P0 = P - npts_h;
  const int Kp = K | 1, Pp = P0 | 1;
  float* s_h = smem + (size_t)TB * (Kp + Pp);            // [TB][npts_h] raw heights (16-byte aligned: Kp + Pp is even... see host)
  __shared__ __align__(8) uint64_t s_hbar;
  __shared__ float s_zref[128];                          // base height - 0.5 of the tile's envs (t1:466)
  float* s_obs = smem + (size_t)tile_warp * 32 * (Kp + Pp);
  float* s_priv = s_obs + 32 * Kp;
  // t1:471-472 observation noise: needs nothing from the other kernels of the step, so a chained launch draws it
  // while ti5_post_physics is still running; it waits in the env's frame slot for the values to be added to it
  __shared__ float s_spawn[SPAWN_FLOATS];
  float org[3] = {0.0f, 0.0f, 0.0f};                     // the env's origin: rewritten only by this env's own reset
  if (do_reset) {
    if (tid == (int)blockDim.x - 1) spawn_state(p, s_spawn);       // an otherwise idle role-1 thread
    if (live) {
#pragma unroll
      for (int i = 0; i < 3; ++i) org[i] = b.env_origins[e * 3 + i];
    }
  }
  // Reset draws ahead of time (Philox mode, small grids): which envs re-spawn is only known after the grid wait, but
  // the draws of a re-spawn depend on (seed, step, env, DOF) alone — the otherwise idle writer warps draw them for
  // EVERY env of the tile while ti5_post_physics runs (26 Philox calls per env, off the critical path) and park them
  // in shared memory; the scatter of a flagged env then only reads them (1.3 -> 0.3 us per flagged env of a warp).
  // (Measured: -0.5 us/step while the SMs were idle in front of the wait; +0.5 us since ti5_post_physics works there in
  // early mode.  Compiled in with -DTI5_PRE_DRAWS only.)
#ifdef TI5_PRE_DRAWS
  const bool pre_draws = do_reset && rng.philox && (int)blockDim.x > OBS_ROLES * TB;
#else
  const bool pre_draws = false;
#endif
  float* s_draw = smem + (size_t)TB * (Kp + Pp + npts_h) + (size_t)(tile_warp * 32 + lane) * DRAW_STRIDE;   // this lane's env
  if (pre_draws && role >= OBS_ROLES && e < N && blockIdx.x < env_blocks) {
    const int half = role - OBS_ROLES;                   // role 2: DOFs 0-5, role 3: DOFs 6-11 and the schedule
#pragma unroll 1
    for (int d = half * (D / 2); d < (half + 1) * (D / 2); ++d) {
      float u[8];
      rng.dof_draws(e, d, u);
#pragma unroll
      for (int i = 0; i < 8; ++i) s_draw[d * 8 + i] = u[i];
    }
    if (half == 1) {
      const float4 a = philox_u4(p.seed, (uint64_t)step, S_LAG, e), c = philox_u4(p.seed, (uint64_t)step, S_GAIT_TIME, e);
      float* sd = s_draw + D * 8;
      sd[0] = a.x; sd[1] = a.y; sd[2] = a.z; sd[3] = a.w; sd[4] = c.x; sd[5] = c.y; sd[6] = c.z; sd[7] = c.w;
    }
  }
  // gait schedule of the env (rewritten only by this env's own reset): for the second command pass (appendix A24)
  int32_t gait_t[TI5_MAX_GAITS] = {0, 0, 0, 0};
  if (do_reset && live) {
#pragma unroll
    for (int gi = 0; gi < TI5_MAX_GAITS; ++gi)
      if (gi < p.num_gaits) gait_t[gi] = b.gait_time[e * p.num_gaits + gi];
  }
  // lagged proprioception rows (pushed up to three steps ago) and the per-env constants only this kernel reads: cold in
  // L2 by now; start fetching them (the lag indices and the stamp only change in this kernel's own reset scatter)
  if (do_obs && live) {
    if (role == 0) {
      const int64_t st0 = b.ring_stamp[e];
      const int64_t jd = (pushes - 1) - b.lag_timestep[e * 3 + 1], ji = (pushes - 1) - b.lag_timestep[e * 3 + 2];
      if ((p.flags & TI5_F_ADD_DOF_LAG) && jd >= st0 && jd >= 0) {
        const float* row = b.dof_ring + ((size_t)ring_slot(jd, p.dof_lag_len) * N + e) * (2 * D);
        prefetch_l2(row);
        prefetch_l2(row + 2 * D - 1);
      }
      if ((p.flags & TI5_F_ADD_IMU_LAG) && ji >= st0 && ji >= 0)
        prefetch_l2(b.imu_ring + ((size_t)ring_slot(ji, p.imu_lag_len) * N + e) * 6);
    } else {
      prefetch_l2(b.env_frictions + e);
      prefetch_l2(b.body_mass + e);
      // simulator rows (evicted since the last step; the kernels of this step do not write them — a pushed or
      // re-spawned root is re-read behind the wait anyway)
      prefetch_l2(b.root_states + (size_t)e * RB);
      prefetch_l2(b.contact_forces + ((size_t)e * NB + p.feet[0]) * 3);
      prefetch_l2(b.contact_forces + ((size_t)e * NB + p.feet[1]) * 3);
    }
  }
  // (dealt over all warps of the tile: while ti5_fused_step fills the register file of an SM, this kernel's CTAs only
  // become resident as its CTAs retire, and whatever stands in front of the wait is then on the critical path — twelve
  // Philox calls on one warp were 1.5 us of it with a warm L2)
  const bool noisy = do_obs && (p.flags & TI5_F_ADD_NOISE);
  if (noisy && e < N && blockIdx.x < env_blocks) {
    float* oo = s_obs + lane * Kp;
    const int nroles = (int)blockDim.x / TB;
#pragma unroll 3
    for (int g4 = role; 4 * g4 < K; g4 += nroles) {
      const float4 u = rng.noise4(e, g4, K);
      const float uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = 4 * g4 + j;
        if (k < K) oo[k] = ((2.0f * uu[j] - 1.0f) * p.noise_vec[k]) * p.noise_level;
      }
    }
  }
  chain_wait();                                          // ti5_post_physics is done
  // t1:466-468: this tile's rows of measured_heights (written by ti5_sample_heights, complete behind the wait): one bulk
  // copy, in flight while the resets and the frames are worked on
  const int h_env0 = blockIdx.x * TB, h_n = min(TB, N - h_env0);
  const bool h_tile = do_obs && npts_h > 0 && blockIdx.x < env_blocks;
  const bool h_bulk = h_tile && ((h_n * npts_h) & 3) == 0 && (((size_t)h_env0 * npts_h) & 3) == 0;
  if (h_bulk && tid == 0) {
    mbar_init(&s_hbar, 1);
    mbar_expect_tx(&s_hbar, (uint32_t)(h_n * npts_h * 4));
    tma_load_1d(s_h, b.measured_heights + (size_t)h_env0 * npts_h, (uint32_t)(h_n * npts_h * 4), &s_hbar);
  }
  if (tid == 0 && atomicAdd(&g->tickets[0], step == INT64_MIN ? 2 : 1) == (int)gridDim.x - 1) {
    g->tickets[0] = 0;
    if (do_obs) g->step_index = step;                    // the step counts as completed: nobody in this grid reads it any more
  }
  int n_reset = 0, id_offset = 0;
  const int64_t counter = step + g->common_step_offset;
  const bool curriculum_due = do_reset && (p.flags & TI5_F_COMMAND_CURRICULUM) && (fast_mod(counter, p.max_episode_length) == 0);
  if (do_reset) {
    int tot = 0, bef = 0;
    double trk = 0.0;
    for (int i = tid; i < env_blocks; i += blockDim.x) {
      const int c = b.block_counts[i];
      tot += c;
      bef += i < (int)blockIdx.x ? c : 0;
      if (curriculum_due && c > 0) trk += (double)b.block_sums[(size_t)i * TI5_LOG_COLS + T_TRACKING_LIN_VEL];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tot += __shfl_xor_sync(0xffffffffu, tot, o);
      bef += __shfl_xor_sync(0xffffffffu, bef, o);
      trk += __shfl_xor_sync(0xffffffffu, trk, o);
    }
    if (lane == 0) { s_tot[warp] = tot; s_bef[warp] = bef; s_trk[warp] = trk; }
    __syncthreads();
    double track_sum = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { n_reset += s_tot[w]; id_offset += s_bef[w]; track_sum += s_trk[w]; }
    if (tid < 6) {
      const int a = tid >> 1, c = tid & 1;
      double v = g->cmd_range[step & 1][a][c];
      if (a == 0 && curriculum_due && n_reset > 0) {      // evaluated before the resets of this step (lr:537-538)
        const float mean = (float)(track_sum / (double)n_reset);
        if (sdiv(mean, (float)p.max_episode_length, dm) > (float)(0.8 * p.tracking_lin_vel_scale)) {
          if (c == 0) { v -= 0.25; const double lo_min = -p.cmd_curriculum_max / 2.0; v = v < lo_min ? lo_min : (v > 0.0 ? 0.0 : v); }
          else { v += 0.5; v = v < 0.0 ? 0.0 : (v > p.cmd_curriculum_max ? p.cmd_curriculum_max : v); }
        }
      }
      s_range[a][c] = v;
      if (blockIdx.x == 0) g->cmd_range[(step + 1) & 1][a][c] = v;     // the next step's copy: nobody reads it in this grid
    }
    __syncthreads();
    if (blockIdx.x == 0) {
      // extras["episode"] snapshot row of this step: previous values if nothing reset (appendix A23)
      float* row = b.extras_log + (size_t)(step % TI5_LOG_ROWS) * TI5_LOG_COLS;
      const float* prev = b.extras_log + (size_t)((step + TI5_LOG_ROWS - 1) % TI5_LOG_ROWS) * TI5_LOG_COLS;
      if (n_reset == 0 && tid < TI5_LOG_COLS) row[tid] = prev[tid];
      if (tid == 0) {
        g->n_reset = n_reset;
        if (n_reset > 0) {
          row[LOG_MAX_COMMAND_X] = (float)s_range[0][1];
          row[LOG_N_RESET] = (float)n_reset;
          if (!(p.flags & TI5_F_TRIMESH)) row[LOG_TERRAIN_LEVEL] = 0.0f;
        }
      }
    }
  } else if (tid < 6) {
    s_range[tid >> 1][tid & 1] = g->cmd_range[step & 1][tid >> 1][tid & 1];
  }
  const bool any_reset = n_reset > 0;
  probe(b.debug_ts, 2, 7);

  const bool flagged = live && do_reset && b.reset_buf[e] != 0;   // both roles see the flag
  const bool reset = flagged && role == 0;                          // ... role 0 accounts for it
  float level_f = 0.0f;
  probe(b.debug_ts, 1, 0);

  // =========================== reset scatter (t1:483-559) ==============================================
  // role 0: the DOF-parallel part, one flagged env at a time with all lanes; role 1 (otherwise idle here): the
  // schedule and the base on the env's own lane.  Flagged envs of a warp go in parallel.
  if (role == 0) {
    unsigned todo = __ballot_sync(0xffffffffu, reset);
    const int env0 = blockIdx.x * TB + tile_warp * 32;
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      reset_env_dofs(p, b, rng, env0 + src, lane, pre_draws ? s_draw + (src - lane) * DRAW_STRIDE : nullptr);
    }
    probe(b.debug_ts, 1, 6);
  } else if (role == 1 && flagged) {
    reset_env_schedule<LAGOPT>(p, b, rng, e, pushes, pre_draws ? s_draw + D * 8 : nullptr);
    reset_env_base(p, b, rng, e, s_spawn, org);
  }
  probe(b.debug_ts, 1, 1);
  __syncthreads();       // role 1 reads what the scatter wrote
  probe(b.debug_ts, 1, 2);

  // The two frames are built in eight parts — obs A1 (command input), A2 (actions + the action copies), A3 (lagged
  // IMU), B1 / B2 (lagged joint positions / velocities), priv A (command input, reference pose, stance), priv B1 (joint
  // state and actions), priv B2 (base, forces, constants) — dealt over the warps the tile has: two (large grids: the
  // two roles take four parts each), four, or eight warps per 32 envs.  A warp runs its part once, on instructions no
  // warp has run before it: the stretch is bound by the length of one warp's cold instruction stream, so more, shorter
  // streams finish earlier as long as the SM has warp slots to spare.
  const int nparts = (int)(blockDim.x / TB);
  const bool builder = e < N && blockIdx.x < env_blocks && role < 8;
  const bool pA1 = role == 0, pPA = role == 1;
  const bool pA2 = nparts >= 8 ? role == 4 : role == 0, pA3 = nparts >= 8 ? role == 5 : role == 0;
  const bool pB1 = nparts >= 4 ? role == 2 : role == 0, pB2 = nparts >= 8 ? role == 6 : pB1;
  const bool pPB1 = nparts >= 4 ? role == 3 : role == 1, pPB2 = nparts >= 8 ? role == 7 : pPB1;
  if (builder) {
    // ======== loads, all issued together ====================================================================
    const bool dof_lag = (p.flags & TI5_F_ADD_DOF_LAG) != 0, imu_lag = (p.flags & TI5_F_ADD_IMU_LAG) != 0;
    float q[D], qd[D];
    if (pPA || pPB1 || ((pB1 || pB2) && !dof_lag)) {
      const float4* ds = reinterpret_cast<const float4*>(b.dof_state + (size_t)e * 2 * D);
#pragma unroll
      for (int i = 0; i < D / 2; ++i) {
        const float4 v = ds[i];
        q[2 * i] = v.x; qd[2 * i] = v.y; q[2 * i + 1] = v.z; qd[2 * i + 1] = v.w;
      }
    }
    float act[D], last_act[D];
    float4 cmd = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t ep_len = 0, phase_len = 0, stamp = 0;
    float gait_start = 0.0f;
    int lag_dof = 0, lag_imu = 0;
    float root[RB], lin[3], ang[3], eul[3], fz0 = 0.0f, fz1 = 0.0f, ext[5], fric = 0.0f, mass = 0.0f;
    if (pA2 || pPB1) load12(b.actions, e, act);
    if (pA1 || pPA) {
      cmd = reinterpret_cast<const float4*>(b.commands)[e];
      ep_len = b.episode_length_buf[e];
      phase_len = b.phase_length_buf[e];
      gait_start = b.gait_start[e];
    }
    if (pA3 || pB1 || pB2) stamp = b.ring_stamp[e];
    if (pB1 || pB2) lag_dof = b.lag_timestep[e * 3 + 1];
    if (pA3) lag_imu = b.lag_timestep[e * 3 + 2];
    if (pA2) load12(b.last_actions, e, last_act);
    if ((pA3 && !imu_lag) || pPB2) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {       // obs A3 needs them as the un-lagged IMU fallback, priv B2 for the critic
        lin[i] = b.base_lin_vel[e * 3 + i];
        ang[i] = b.base_ang_vel[e * 3 + i];
        eul[i] = b.base_euler_xyz[e * 3 + i];
      }
    }
    if (pPB2) {
#pragma unroll
      for (int i = 0; i < RB; ++i) root[i] = b.root_states[(size_t)e * RB + i];
      fz0 = b.contact_forces[((size_t)e * NB + p.feet[0]) * 3 + 2];
      fz1 = b.contact_forces[((size_t)e * NB + p.feet[1]) * 3 + 2];
      const bool xf = p.flags & TI5_F_ADD_EXT_FORCE;
      ext[0] = (xf ? b.ext_forces : b.rand_push_force)[e * 3 + 0];
      ext[1] = (xf ? b.ext_forces : b.rand_push_force)[e * 3 + 1];
#pragma unroll
      for (int i = 0; i < 3; ++i) ext[2 + i] = (xf ? b.ext_torques : b.rand_push_torque)[e * 3 + i];
      fric = b.env_frictions[e];
      mass = b.body_mass[e];
    }

    // t1:527 `_resample_commands()` runs over ALL envs whenever anything reset (appendix A24)
    if (do_reset && any_reset && (pA1 || pPA)) {
      for (int gi = 0; gi < p.num_gaits; ++gi) {
        const int32_t gt = flagged ? b.gait_time[e * p.num_gaits + gi] : gait_t[gi];     // a re-spawned env has a new schedule
        if (ep_len != (int64_t)gt) continue;
        const int kind = p.gait_kind[gi];
        const bool mx = kind == TI5_GAIT_WALK_SAGITTAL || kind == TI5_GAIT_WALK_OMNI;
        const bool my = kind == TI5_GAIT_WALK_LATERAL || kind == TI5_GAIT_WALK_OMNI;
        const bool mz = kind == TI5_GAIT_ROTATE || kind == TI5_GAIT_WALK_OMNI;
        float cu[3];
        rng.cmd3(1, gi, e, N, cu);
        cmd.x = mx ? affine((float)(s_range[0][1] - s_range[0][0]), (float)s_range[0][0], cu[0]) : 0.0f;
        cmd.y = my ? affine((float)(s_range[1][1] - s_range[1][0]), (float)s_range[1][0], cu[1]) : 0.0f;
        // heading mode: the heading target is re-drawn; the yaw rate keeps the value of this step's callback (t1:185-188
        // runs before the resets)
        if (TI5_HEADING(p)) cmd.w = mz ? affine(p.heading_w, p.heading_lo, cu[2]) : 0.0f;
        else cmd.z = mz ? affine((float)(s_range[2][1] - s_range[2][0]), (float)s_range[2][0], cu[2]) : 0.0f;
      }
      if (pA1) {
        reinterpret_cast<float4*>(b.commands)[e] = cmd;
        if (b.time_outs_latched) b.time_outs_latched[e] = b.time_out_buf[e];    // t1:540-541 (appendix A23)
      }
    }
    if (role == 0 && (p.flags & TI5_F_TRIMESH)) level_f = (float)b.terrain_levels[e];
    probe(b.debug_ts, 2, 2);        // (builder probes: kernel row 2, slots 2-7; obs A1 = thread 0)

    // =========================== observations (t1:368-481) =================================
    if (do_obs) {
      float s = 0.0f, ci[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      if (pA1 || pPA) {
        const bool no_sw = (p.flags & TI5_F_NO_SW_SWITCH) != 0;                 // t1:89-90: phase from the episode counter
        const bool stand = !no_sw && sqrtf(cmd.x * cmd.x + cmd.y * cmd.y + cmd.z * cmd.z) <= p.stand_threshold;
        if (stand) phase_len = 0;                                               // t1:86 side effect
        const float phase = (py_mod1(sdiv((float)(no_sw ? ep_len : phase_len) * p.dt, p.cycle_time, dm)) + gait_start) * (stand ? 0.0f : 1.0f);
        const float ang_ph = TWO_PI_F * phase;
        s = sinf(ang_ph);
        ci[0] = s; ci[1] = cosf(ang_ph);
        ci[2] = cmd.x * p.cmd_scale[0]; ci[3] = cmd.y * p.cmd_scale[1]; ci[4] = cmd.z * p.cmd_scale[2];
      }
      float* oo = s_obs + lane * Kp;
      float* po = s_priv + lane * Pp;
      auto put = [&](int k, float v) { oo[k] = noisy ? v + oo[k] : v; };       // value + noise (drawn above)
      // per-step re-draw of a lag index: never more than one step further back than the last one (t1:411-412)
      auto relag = [&](int kind) {
        const int draw = rng.lag_step_draw(kind, e, N);
        const int last = b.last_lag[((size_t)(step & 1) * N + e) * 5 + kind];
        return draw > last + 1 ? last + 1 : draw;
      };

      if (pA1) {
        // ---------------- obs A1: command input (t1:407-411) -------------------------------------------------
        b.phase_length_buf[e] = phase_len;
#pragma unroll
        for (int i = 0; i < 5; ++i) put(i, ci[i]);
      }
      if (pA2) {
        // ---------------- obs A2: actions; lr:496-497 previous-step action copies ---------------------------
#pragma unroll
        for (int i = 0; i < D; ++i) put(29 + i, act[i]);
        store12(b.last_last_actions, e, last_act);
        store12(b.last_actions, e, act);
      }
      if (pA3) {
        // ---------------- obs A3: lagged IMU (t1:438-451) ---------------------------------------------------
        float imu[6];
        if (LAGOPT && (p.flags & TI5_F_IMU_LAG_PERSTEP)) {   // t1:437-442
          lag_imu = relag(2);
          b.lag_timestep[e * 3 + 2] = lag_imu;
          b.last_lag[((size_t)((step + 1) & 1) * N + e) * 5 + 2] = lag_imu;
        }
        const int64_t ji = (pushes - 1) - lag_imu;
        if (!imu_lag) {
#pragma unroll
          for (int i = 0; i < 3; ++i) { imu[i] = ang[i]; imu[3 + i] = eul[i]; }
        } else if (ji >= stamp && ji >= 0) {
          const float* row = b.imu_ring + ((size_t)ring_slot(ji, p.imu_lag_len) * N + e) * 6;
#pragma unroll
          for (int i = 0; i < 6; ++i) imu[i] = row[i];
        } else {
#pragma unroll
          for (int i = 0; i < 6; ++i) imu[i] = 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          put(41 + i, imu[i] * p.obs_ang_vel);
          put(44 + i, imu[3 + i] * p.obs_quat);
        }
        probe(b.debug_ts, 2, 3, nparts >= 8 ? 5 * TB : 0);
      }
      if (pB1 || pB2) {
        // ---------------- obs B1 / B2: lagged joint positions / velocities; rows pushed before the env's last reset
        // read as zero ----------------------------------------------------------------------------------------
        // the lag of the positions (B1) and of the velocities (B2): the common joint-state lag, unless the options t1_cfg
        // leaves off are in play — a per-step re-draw (t1:408-413), or separate position / velocity lags (t1:416-431)
        int lag_q = lag_dof, lag_qd = lag_dof;
        if (LAGOPT && (p.flags & TI5_F_LAG_OPTIONS)) {
          const size_t nxt = ((size_t)((step + 1) & 1) * N + e) * 5;
          if (p.flags & TI5_F_POS_VEL_LAG) {
            lag_q = b.lag_pv[e * 2 + 0];
            lag_qd = b.lag_pv[e * 2 + 1];
            if (pB1 && (p.flags & TI5_F_POS_LAG_PERSTEP)) { lag_q = relag(3); b.lag_pv[e * 2 + 0] = lag_q; b.last_lag[nxt + 3] = lag_q; }
            if (pB2 && (p.flags & TI5_F_VEL_LAG_PERSTEP)) { lag_qd = relag(4); b.lag_pv[e * 2 + 1] = lag_qd; b.last_lag[nxt + 4] = lag_qd; }
          } else if (p.flags & TI5_F_DOF_LAG_PERSTEP) {
            lag_q = lag_qd = relag(1);
            if (pB1) { b.lag_timestep[e * 3 + 1] = lag_q; b.last_lag[nxt + 1] = lag_q; }
          }
        }
        const int64_t jj = (pushes - 1) - lag_q;
        const bool hit = jj >= stamp && jj >= 0;
        const float* row = hit ? b.dof_ring + ((size_t)ring_slot(jj, p.dof_lag_len) * N + e) * (2 * D) : nullptr;
        bool hit2 = hit;
        const float* row2 = row;
        if (LAGOPT && lag_qd != lag_q) {                       // separate velocity lag only
          const int64_t jj2 = (pushes - 1) - lag_qd;
          hit2 = jj2 >= stamp && jj2 >= 0;
          row2 = hit2 ? b.dof_ring + ((size_t)ring_slot(jj2, p.dof_lag_len) * N + e) * (2 * D) : nullptr;
        }
        if (pB1) {
          float lq[D];
          if (!dof_lag) {
#pragma unroll
            for (int i = 0; i < D; ++i) lq[i] = q[i];
          } else if (hit) {
            load12(row, 0, lq);
          } else {
#pragma unroll
            for (int i = 0; i < D; ++i) lq[i] = 0.0f;
          }
#pragma unroll
          for (int i = 0; i < D; ++i) put(5 + i, (lq[i] - p.default_dof_pos[i]) * p.obs_dof_pos);
        }
        if (pB2) {
          float lqd[D];
          if (!dof_lag) {
#pragma unroll
            for (int i = 0; i < D; ++i) lqd[i] = qd[i];
          } else if (hit2) {
            load12(row2 + D, 0, lqd);
          } else {
#pragma unroll
            for (int i = 0; i < D; ++i) lqd[i] = 0.0f;
          }
#pragma unroll
          for (int i = 0; i < D; ++i) put(17 + i, lqd[i] * p.obs_dof_vel);
          probe(b.debug_ts, 2, 4, nparts >= 8 ? 6 * TB : (nparts >= 4 ? 2 * TB : 0));
        }
      }
      if (pPA) {
        // ---------------- priv A: command input, reference pose (t1:250-274), stance -------------------------
        float ref[D];
#pragma unroll
        for (int i = 0; i < D; ++i) ref[i] = 0.0f;
        {
          const float sl = s > 0.0f ? 0.0f : s, sr = s < 0.0f ? 0.0f : s;
          const float a1 = p.target_joint_pos_scale, a2 = p.target_joint_pos_scale2;
          ref[2] = sl * a1; ref[3] = (-sl) * a2; ref[4] = sl * a1;
          ref[8] = (-sr) * a1; ref[9] = sr * a2; ref[10] = (-sr) * a1;
          if (fabsf(s) < 0.1f) {
#pragma unroll
            for (int i = 0; i < D; ++i) ref[i] = 0.0f;
          }
        }
        float ref_act[D];
#pragma unroll
        for (int i = 0; i < D; ++i) { ref_act[i] = 2.0f * ref[i]; ref[i] = ref[i] + p.default_dof_pos[i]; }
        store12(b.ref_action, e, ref_act);
        store12(b.ref_dof_pos, e, ref);
        float stance[2] = {s >= 0.0f ? 1.0f : 0.0f, s < 0.0f ? 1.0f : 0.0f};
        if (fabsf(s) < 0.1f) stance[0] = stance[1] = 1.0f;
#pragma unroll
        for (int i = 0; i < 5; ++i) po[i] = ci[i];
#pragma unroll
        for (int i = 0; i < D; ++i) po[41 + i] = q[i] - ref[i];
        po[69] = stance[0]; po[70] = stance[1];
        probe(b.debug_ts, 2, 5, TB);
      }
      if (pPB1) {
        // ---------------- priv B1: joint state and actions; lr:498 previous-step joint velocities --------------
#pragma unroll
        for (int i = 0; i < D; ++i) {
          po[5 + i] = (q[i] - p.default_dof_pos[i]) * p.obs_dof_pos;
          po[17 + i] = qd[i] * p.obs_dof_vel;
          po[29 + i] = act[i];
        }
        store12(b.last_dof_vel, e, qd);
      }
      if (pPB2) {
        // ---------------- priv B2: base state, disturbance forces, per-env constants, contacts ------------------
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          po[53 + i] = lin[i] * p.obs_lin_vel;
          po[56 + i] = ang[i] * p.obs_ang_vel;
          po[59 + i] = eul[i] * p.obs_quat;
        }
        if (p.flags & TI5_F_ADD_EXT_FORCE) {                                   // t1:386-388
          po[62] = sdiv(ext[0], p.ext_force_div, dm);
          po[63] = sdiv(ext[1], p.ext_force_div, dm);
#pragma unroll
          for (int i = 0; i < 3; ++i) po[64 + i] = sdiv(ext[2 + i], p.ext_torque_div, dm);
        } else {
#pragma unroll
          for (int i = 0; i < 5; ++i) po[62 + i] = ext[i];
        }
        po[67] = fric;
        po[68] = sdiv(mass, 30.0f, dm);
        po[71] = fz0 > 5.0f ? 1.0f : 0.0f; po[72] = fz1 > 5.0f ? 1.0f : 0.0f;
        if (npts_h) s_zref[le] = root[2] - 0.5f;                               // t1:466-468: used by the ring writers
#pragma unroll
        for (int i = 0; i < 6; ++i) b.last_root_vel[e * 6 + i] = root[7 + i];  // lr:499
        probe(b.debug_ts, 2, 6, nparts >= 8 ? 7 * TB : (nparts >= 4 ? 3 * TB : TB));
      }
    }
  }

  // partial last tile whose byte count the bulk engine does not take: the tile's threads copy the raw heights
  if (h_tile && !h_bulk) {
    for (int i = tid; i < h_n * npts_h; i += blockDim.x) s_h[i] = b.measured_heights[(size_t)h_env0 * npts_h + i];
  }
  probe(b.debug_ts, 1, 3);
  // ---- history rings: append this step's frames; clear the rows of re-spawned envs ------------
  const size_t obs_row = (size_t)2 * H * K, priv_row = (size_t)2 * CH * P;
  const int warp_env0 = blockIdx.x * TB + tile_warp * 32;
  const int hs = (int)fast_mod(step - 1, H), cs = (int)fast_mod(step - 1, CH);       // slot of this step's frame
  __syncthreads();       // both frames of every env of the CTA are staged
  if (h_bulk) mbar_wait(&s_hbar, 0);                     // ... and the tile's raw heights have landed
  probe(b.debug_ts, 1, 7);
  if (do_obs && blockIdx.x < env_blocks && warp_env0 < N) {
    const int n_here = min(32, N - warp_env0);
    const float lim = p.clip_obs;
    const int L = p.log_len, ls = L > 0 ? (int)fast_mod(step - 1, L) : 0;      // frame-log row of this step
    // the tile's warps (frame builders and writers alike) stride over the staged obs elements, then the priv ones;
    // small loop bodies on purpose (32-bit offsets from a uniform base, no unrolling).  (A lane-per-element / loop-over-
    // envs form with half the instructions measured no faster at 8192 envs and 25 % slower at 65536 with the 260-float
    // privileged frame: the stretch is bound by the store path, and this order spreads consecutive stores over more rows.)
    const int n_obs = n_here * K, n_priv = n_here * P, stride = (int)(blockDim.x / TB) * 32;
    const uint32_t orow = (uint32_t)obs_row, prow = (uint32_t)priv_row, omir = (uint32_t)(H * K), pmir = (uint32_t)(CH * P);
    float* ob = b.obs_ring + (size_t)warp_env0 * obs_row + (size_t)hs * K;
    float* pb = b.priv_ring + (size_t)warp_env0 * priv_row + (size_t)cs * P;
    float* ol = L > 0 ? b.frame_log + ((size_t)warp_env0 * L + ls) * K : nullptr;
    float* pl = L > 0 ? b.priv_log + ((size_t)warp_env0 * L + ls) * P : nullptr;
#pragma unroll 1
    for (int i = role * 32 + lane; i < n_obs; i += stride) {
      const int en = i / K, k = i - en * K;
      const float v = clampf(s_obs[en * Kp + k], -lim, lim);
      float* dst = ob + (uint32_t)en * orow + (uint32_t)k;
      dst[0] = v;
      dst[omir] = v;
      if (L > 0) ol[(uint32_t)en * (uint32_t)(L * K) + (uint32_t)k] = v;
    }
    auto priv_entry = [&](int en, int k) {      // entry k of env en's privileged frame, clipped
      float raw;
      if (k < P0) raw = s_priv[en * Pp + k];
      else raw = clampf(s_zref[tile_warp * 32 + en] - s_h[(tile_warp * 32 + en) * npts_h + (k - P0)], -1.0f, 1.0f) * p.obs_height;
      return clampf(raw, -lim, lim);
    };
    if (PC > 0 && (PC & 3) == 0) {
      // a frame width that is a multiple of four floats (260 with measured heights) makes every frame of the critic ring
      // start 16-byte aligned (row 2*CH*P, slot P, mirror CH*P, log L*P floats): one 128-bit store per four entries
      constexpr int P4 = PC > 0 ? PC / 4 : 1;
#pragma unroll 1
      for (int i = role * 32 + lane; i < n_here * P4; i += stride) {
        const int en = i / P4, k = (i - en * P4) * 4;
        const float4 v = make_float4(priv_entry(en, k), priv_entry(en, k + 1), priv_entry(en, k + 2), priv_entry(en, k + 3));
        float* dst = pb + (uint32_t)en * prow + (uint32_t)k;
        *reinterpret_cast<float4*>(dst) = v;
        *reinterpret_cast<float4*>(dst + pmir) = v;
        if (L > 0) *reinterpret_cast<float4*>(pl + (uint32_t)en * (uint32_t)(L * P) + (uint32_t)k) = v;
      }
    } else {
#pragma unroll 1
      for (int i = role * 32 + lane; i < n_priv; i += stride) {
        const int en = i / P, k = i - en * P;
        const float v = priv_entry(en, k);
        float* dst = pb + (uint32_t)en * prow + (uint32_t)k;
        dst[0] = v;
        dst[pmir] = v;
        if (L > 0) pl[(uint32_t)en * (uint32_t)(L * P) + (uint32_t)k] = v;
      }
    }
  }
  // frames of the env's window that no reset has cleared: 0 after reset_idx, +1 per appended frame
  if (p.log_len > 0 && role == 0 && live) {
    int hv = b.hist_valid[e];
    if (do_reset && reset) hv = 0;
    if (do_obs) {
      if (s_new_step) hv = min(hv + 1, H);      // a re-observation rewrites the slot of the step: no frame is added
      b.valid_log[(size_t)fast_mod(step - 1, p.log_len) * N + e] = (int16_t)hv;
    }
    b.hist_valid[e] = hv;
  }

  probe(b.debug_ts, 1, 4);
  // t1:556-559: `hist[i][env_ids] *= 0` for every frame of the re-spawned envs.  The flagged envs were
  // listed (in arrival order) by ti5_post_physics / ti5_reset_bookkeeping; all warps of the grid,
  // including the helper CTAs, share the (env, 512-float chunk) work items.  The two slots of this step's
  // frame are skipped — their owner overwrites them above — so no ordering between warps is needed.
  const int n_list = do_reset ? g->n_listed[step & 1] : 0;
  if (n_list > 0) {
    constexpr int CHUNK = 512, U = CHUNK / 32;
    const int oc = (int)((obs_row + CHUNK - 1) / CHUNK), pc = (int)((priv_row + CHUNK - 1) / CHUNK);
    const int per_env = oc + pc;
    const int wpb = blockDim.x >> 5;
    const int total = n_list * per_env;
    // The helper CTAs start on the list at once; the env CTAs get here only after their frames are out.  With the
    // usual handful of re-spawned envs the helpers finish long before that, so they take the whole list and the
    // clear leaves the env CTAs' critical path; a mass reset (reset() of every env) is shared by all warps.
    const int helper_warps = ((int)gridDim.x - env_blocks) * wpb;
    const bool helpers_only = helper_warps > 0 && total <= 8 * helper_warps;
    int w0, stride;
    if (helpers_only) {
      w0 = blockIdx.x >= env_blocks ? (blockIdx.x - env_blocks) * wpb + warp : total;
      stride = helper_warps;
    } else {
      w0 = blockIdx.x * wpb + warp;
      stride = gridDim.x * wpb;
    }
    for (int w = w0; w < total; w += stride) {
      const int en = b.reset_list[w / per_env];
      int c = w % per_env;
      const bool is_obs = c < oc;
      if (!is_obs) c -= oc;
      float* row = is_obs ? b.obs_ring + (size_t)en * obs_row : b.priv_ring + (size_t)en * priv_row;
      const int len = (int)(is_obs ? obs_row : priv_row), width = is_obs ? K : P;
      const int s0 = (is_obs ? hs : cs) * width, s1 = (is_obs ? hs + H : cs + CH) * width;
      // `x *= 0` is stored as +0 without reading x back (the product of a finite x is +-0, equal to it in every
      // comparison and in every sum the policy forms; the read-modify-write cost a DRAM round trip per chunk on the
      // critical path of every CTA of a large grid: 7.4 us per CTA at 65536 envs).  A chunk clear of this step's two
      // frame slots, in a row whose length is a multiple of four floats (every chunk then starts 16-byte aligned), is
      // four 128-bit stores per lane; the chunks that hold the frame slots go element by element.
      const int c0 = c * CHUNK, c1 = min(c0 + CHUNK, len);
      const bool touches_frame = do_obs && ((c0 < s0 + width && c1 > s0) || (c0 < s1 + width && c1 > s1));
      if ((len & 3) == 0 && !touches_frame) {
        float4* r4 = reinterpret_cast<float4*>(row + c0);
        const int n4 = (c1 - c0) >> 2;
#pragma unroll
        for (int j = 0; j < CHUNK / 4 / 32; ++j)
          if (j * 32 + lane < n4) r4[j * 32 + lane] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        continue;
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int i = c0 + j * 32 + lane;
        if (i < len && !(do_obs && ((i >= s0 && i < s0 + width) || (i >= s1 && i < s1 + width)))) row[i] = 0.0f;
      }
    }
  }
  probe(b.debug_ts, 1, 5);
  if (blockIdx.x >= env_blocks) {             // helper CTA
    // extras["episode"]["rew_<term>"] = mean over the re-spawned envs of the episode sums / episode_length_s
    // (t1:531-532): one helper warp per term folds the per-CTA partials ti5_post_physics left behind.
    const int hw = (blockIdx.x - env_blocks) * (blockDim.x >> 5) + warp;
    const int total = n_reset;
    if (do_reset && total > 0 && hw < TI5_NUM_TERMS) {
      double acc = 0.0;
      for (int i = lane; i < env_blocks; i += 32) {
        const int cnt = b.block_counts[i];
        const float v = b.block_sums[(size_t)i * TI5_LOG_COLS + hw];
        acc += cnt > 0 ? (double)v : 0.0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        const float mean = (float)(acc / (double)total);
        b.extras_log[(size_t)(step % TI5_LOG_ROWS) * TI5_LOG_COLS + hw] = sdiv(mean, p.max_episode_length_s, dm);
      }
    }
    return;
  }

  // host-side callers: this step's scalar outputs, stored straight into (mapped, pinned) host memory
  if (b.host_out && do_obs && role == 0 && live) {
    reinterpret_cast<float*>(b.host_out)[e] = b.rew_buf[e];
    b.host_out[(size_t)4 * N + e] = b.reset_buf[e];
    b.host_out[(size_t)5 * N + e] = b.time_outs_latched ? b.time_outs_latched[e] : 0;
  }

  // ---- ascending id list of the envs reset this step (lr:490) --------------------------------
  const BlockRank br = block_rank(reset, s_warp);
  if (reset && b.reset_ids) b.reset_ids[id_offset + br.rank] = e;
  if (reset && b.dof_props) {       // lr:915-939: the re-drawn joint properties, dense, in the order of the id list
    float* o = b.dof_props + (size_t)(id_offset + br.rank) * D * 3;
    const float* arm = b.joint_armatures + (size_t)e * D;      // written by this CTA's scatter, before a __syncthreads
#pragma unroll 1
    for (int d = 0; d < D; ++d) {
      // (the env's friction / damping multipliers were stored by role 1 of this CTA in front of a __syncthreads)
      o[d * 3 + 0] = (LAGOPT && (p.flags2 & TI5_F2_RAND_JOINT_FRICTION)) ? b.joint_coeffs[e * 2 + 0] : 1.0f;
      o[d * 3 + 1] = (LAGOPT && (p.flags2 & TI5_F2_RAND_JOINT_DAMPING)) ? b.joint_coeffs[e * 2 + 1] : 1.0f;
      o[d * 3 + 2] = (p.flags & TI5_F_RAND_ARMATURE) ? arm[d] : 0.0f;
    }
  }

  // ---- extras["episode"]["terrain_level"] = mean(terrain_levels) (t1:535-536) ------------------
  if ((p.flags & TI5_F_TRIMESH) && do_reset) {
    float v = level_f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_lvl[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.0f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_lvl[w];
      b.block_sums[(size_t)blockIdx.x * TI5_LOG_COLS + LOG_TERRAIN_LEVEL] = t;
      __threadfence();
      s_last = atomicAdd(&g->tickets[1], 1) == env_blocks - 1;
    }
    __syncthreads();
    if (s_last) {                                  // whole-CTA fold of the per-CTA level sums
      __threadfence();
      float part = 0.0f;
      for (int i = threadIdx.x; i < env_blocks; i += blockDim.x)
        part += __ldcg(b.block_sums + (size_t)i * TI5_LOG_COLS + LOG_TERRAIN_LEVEL);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      __syncthreads();
      if (lane == 0) s_lvl[warp] = part;
      __syncthreads();
      if (threadIdx.x == 0) {
        g->tickets[1] = 0;
        if (any_reset) {
          double acc = 0.0;
          for (int w = 0; w < (int)(blockDim.x >> 5); ++w) acc += (double)s_lvl[w];
          b.extras_log[(size_t)(step % TI5_LOG_ROWS) * TI5_LOG_COLS + LOG_TERRAIN_LEVEL] = (float)(acc / (double)N);
        }
      }
    }
  }
}

// lr:441-446 / t1:477-481: contiguous copies of the current windows for callers that need them
__global__ void __launch_bounds__(256) materialize_kernel(const __grid_constant__ Ti5Params p, const __grid_constant__ Ti5Buffers b) {
  const int N = p.num_envs, K = p.num_single_obs, P = p.priv_frame, H = p.frame_stack, CH = p.c_frame_stack;
  const int64_t step = b.globals->step_index;            // runs after the step has been completed
  const size_t ow = (size_t)H * K, pw = (size_t)CH * P;
  const size_t o_off = (size_t)(fast_mod(step - 1, H) + 1) * K, p_off = (size_t)(fast_mod(step - 1, CH) + 1) * P;
  const size_t total_o = (size_t)N * ow, total_p = (size_t)N * pw;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_o + total_p; i += stride) {
    if (i < total_o) {
      if (b.obs_out) {
        const size_t en = i / ow, k = i - en * ow;
        b.obs_out[i] = b.obs_ring[en * 2 * ow + o_off + k];
      }
    } else if (b.priv_out) {
      const size_t j = i - total_o, en = j / pw, k = j - en * pw;
      b.priv_out[j] = b.priv_ring[en * 2 * pw + p_off + k];
    }
  }
}

}  // namespace ti5

using namespace ti5;

extern "C" int ti5_reset_observe(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, int phases, void* stream) {
  TI5_CHECK_ARGS(p && b && p->num_envs > 0 && (phases & 3) != 0 && (phases & ~7) == 0);
  // a chained launch takes the step index from the completed-step counter: only valid for the full step
  TI5_CHECK_ARGS(!(phases & TI5_RO_CHAINED) || (phases & 3) == 3);
  TI5_CHECK_ARGS(p->env_block == 32 || p->env_block == 64 || p->env_block == 128);
  TI5_CHECK_ARGS(p->num_single_obs == 47 && p->priv_frame >= 73 && p->num_single_obs <= 64);
  TI5_CHECK_ARGS(p->rng_mode == TI5_RNG_PHILOX || (r && r->cmd && r->dofs && r->dr && r->gait_time && r->noise));
  {   // the lag options t1_cfg leaves off: their state and (parity mode) their draws
    const int perstep = TI5_F_LAG_PERSTEP | TI5_F_DOF_LAG_PERSTEP | TI5_F_IMU_LAG_PERSTEP | TI5_F_POS_LAG_PERSTEP | TI5_F_VEL_LAG_PERSTEP;
    TI5_CHECK_ARGS(!(p->flags & perstep) || b->last_lag);
    TI5_CHECK_ARGS(!(p->flags & (perstep & ~TI5_F_LAG_PERSTEP)) || p->rng_mode == TI5_RNG_PHILOX || (r && r->lag_step));
    TI5_CHECK_ARGS(!(p->flags & TI5_F_POS_VEL_LAG) || (b->lag_pv && (p->flags & TI5_F_ADD_DOF_LAG) &&
                                                        (p->rng_mode == TI5_RNG_PHILOX || (r && r->lag_idx_pv))));
  }
  Ti5Rng rr = r ? *r : Ti5Rng{};
  const int blocks = (p->num_envs + p->env_block - 1) / p->env_block;
  // warps per 32 envs (2, 4 or 8; TI5_RO_ROLES overrides): eight on the small grids, where the SMs have warp slots to
  // spare and the frame-building stretch is bound by the length of one warp's instruction stream; two on large grids
  // (the SMs are full of frame builders there; more warps per env would only take their registers)
  static const int forced_roles = getenv("TI5_RO_ROLES") ? atoi(getenv("TI5_RO_ROLES")) : 0;
  // ... and four on large grids WITH measured heights: a tile of 64 envs stages 79 KB there, so only two CTAs fit an SM,
  // and two roles per env would leave it with eight warps (65536 envs, config 3: this kernel 187 -> 156 us; without the
  // heights four CTAs of 128 threads fit and four roles measured 5 us slower than two)
  const bool heights_staged = (p->flags & TI5_F_MEASURE_HEIGHTS) && p->num_height_points > 0;
  int nroles = forced_roles ? forced_roles : (p->env_block == 32 ? 8 : (p->env_block == 64 && heights_staged) ? 4 : 2);
  while (nroles > 2 && nroles * p->env_block > 256) nroles >>= 1;
  TI5_CHECK_ARGS(nroles == 2 || nroles == 4 || nroles == 8);
  const bool writers = nroles > 2;
#ifdef TI5_PRE_DRAWS
  const int draw_floats = writers ? DRAW_STRIDE : 0;
#else
  const int draw_floats = 0;
#endif
  const int npts_h = (p->flags & TI5_F_MEASURE_HEIGHTS) ? p->num_height_points : 0;
  TI5_CHECK_ARGS(npts_h == 0 || p->priv_frame - npts_h == 73);
  // per env: 47 (+0: odd) observation floats, 73 privileged floats staged per env, the raw heights of the tile behind them
  // ((47 | 1) + (73 | 1) = 120 floats = 480 bytes per env: the heights block starts 16-byte aligned)
  const size_t smem = (size_t)p->env_block * ((p->num_single_obs | 1) + ((p->priv_frame - npts_h) | 1) + npts_h + draw_floats) * sizeof(float);
  // the 128-register build: four CTAs of 128 threads per SM on large grids, two CTAs of 256 threads with eight roles
  const bool big = p->num_envs >= 32768 || nroles * p->env_block > 128;
  auto kernel = p->priv_frame == 73 ? (big ? reset_observe_kernel<47, 73, 2> : reset_observe_kernel<47, 73, 1>)
                : p->priv_frame == 260 ? (big ? reset_observe_kernel<47, 260, 2> : reset_observe_kernel<47, 260, 1>)
                                       : (big ? reset_observe_kernel<47, 0, 2> : reset_observe_kernel<47, 0, 1>);
  TI5_CHECK_ARGS(p->flags2 == 0 || (b->joint_coeffs && (p->rng_mode == TI5_RNG_PHILOX || (r && r->dr_joint))));
  if ((p->flags & TI5_F_LAG_OPTIONS) || p->flags2)       // the rarely used options: the generic-width builds carry them
    kernel = big ? reset_observe_kernel<47, 0, 2, true> : reset_observe_kernel<47, 0, 1, true>;
  if (!ti5_ensure_smem(kernel, smem)) {
    ti5_set_error("ti5_reset_observe: %zu bytes of shared memory per CTA not available", smem);
    return TI5_ECUDA;
  }
  ti5_set_carveout(kernel, ti5_small_grid(p));
  // + helper CTAs (about 4 warps per SM) that only share the history-clear work of re-spawned envs
  // writer warps only for the small-grid case (env_block 32): on larger grids the SMs are full of frame builders and
  // idle writers would only take their registers
  const int threads = nroles * p->env_block;   // <= 256
  // a grid that fits the GPU in one wave without the helpers keeps it that way with them: the CTAs that would start in a
  // second wave were 0.8 us per step at 8192 envs with eight roles (256 + 74 CTAs on 296 slots)
  int helpers = (phases & TI5_RO_RESET) ? (ti5_sm_count() * 4 * 32) / threads : 0;
  if (helpers > 0) {
    const int room = ti5_sm_count() * ti5_ctas_per_sm(kernel, threads, smem) - blocks;
    if (room >= 16 && room < helpers) helpers = room;
  }
  (void)ti5_launch(kernel, dim3(blocks + helpers), dim3(threads), smem, stream, (phases & TI5_RO_CHAINED) != 0, *p, *b, rr, phases);
  return ti5_check_launch("ti5_reset_observe");
}

extern "C" int ti5_reset_scatter(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, void* stream) {
  return ti5_reset_observe(p, b, r, TI5_RO_RESET, stream);
}

extern "C" int ti5_observations(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, void* stream) {
  return ti5_reset_observe(p, b, r, TI5_RO_OBSERVE, stream);
}

extern "C" int ti5_materialize_obs(const Ti5Params* p, const Ti5Buffers* b, void* stream) {
  TI5_CHECK_ARGS(p && b && p->num_envs > 0 && (b->obs_out || b->priv_out));
  materialize_kernel<<<ti5_sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(*p, *b);
  return ti5_check_launch("ti5_materialize_obs");
}
