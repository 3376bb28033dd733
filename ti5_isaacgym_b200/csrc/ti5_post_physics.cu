// Post-physics phase, part 1 (once per policy step): counters, derived base state, command
// schedule, push / external-force windows, termination, the reward sum, and the reset
// bookkeeping.  One thread per env; the AoS simulator rows are read once into registers and
// every (TERMS, N) / (N, k) product array is written coalesced.
//
// Replaces, in the reference's order:  lr:464-481 (post_physics_step head), t1:179-215
// (_post_physics_step_callback), lr:509-517 (check_termination), lr:654-680 + t1:572-946
// (compute_reward and the reward terms), and the reductions reset_idx needs before it can run
// (lr:490 count, t1:530-541 episode means, lr:1160-1169 command curriculum).
#include "ti5_device.cuh"
#include "ti5_host.h"

namespace ti5 {

struct FootState {
  float pos[3];
  float quat[4];
  float wxy[2];     // rigid_state[..., 10:12] (appendix A10)
  float force[3];
  float pitch;      // feet_euler_xyz[..., 1]
};

// t1:599-628 feet_distance / knee_distance
__device__ __forceinline__ float pair_distance_reward(float ax, float ay, float bx, float by, float lo, float hi) {
  const float dx = ax - bx, dy = ay - by;
  const float d = sqrtf(dx * dx + dy * dy);
  const float near_ = clampf(d - lo, -0.5f, 0.0f);
  const float far_ = clampf(d - hi, 0.0f, 0.5f);
  return (expf(-fabsf(near_) * 100.0f) + expf(-fabsf(far_) * 100.0f)) / 2.0f;
}

// Reset bookkeeping shared by ti5_post_physics and ti5_reset_bookkeeping: per-CTA reset counts
// and episode-sum partials; the last CTA to finish turns the counts into exclusive offsets
// (consumed by ti5_reset_observe for the ascending id list), publishes n_reset, writes the
// extras["episode"] snapshot row of this step and evaluates the command curriculum.
__device__ __forceinline__ void reset_bookkeeping(const Ti5Params& p, const Ti5Buffers& b, bool reset,
                                                  const float (&esum)[TI5_NUM_TERMS], int64_t step, int64_t counter,
                                                  bool force_window, bool advance_force_flag) {
  __shared__ int s_warp[32];
  __shared__ float s_red[4][TI5_NUM_TERMS];
  __shared__ bool s_last;
  Ti5Globals* g = b.globals;
  const int dm = p.div_mode;
  // ---- reset bookkeeping: per-CTA count and episode-sum partials; the last CTA finishes ------
  const BlockRank br = block_rank(reset, s_warp);
  const int nblk = gridDim.x;
  if (br.total > 0) {
    // sum of the resetting envs' episode sums (t1:531-533), one column per term
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < TI5_NUM_TERMS; ++t) {
      float v = reset ? esum[t] : 0.0f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_red[warp][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < TI5_NUM_TERMS) {
      float v = 0.0f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += s_red[w][threadIdx.x];
      b.block_sums[(size_t)blockIdx.x * TI5_LOG_COLS + threadIdx.x] = v;
    }
  }
  if (threadIdx.x == 0) {
    b.block_counts[blockIdx.x] = br.total;
    __threadfence();
    s_last = atomicAdd(&g->tickets[0], 1) == nblk - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // exclusive prefix of the CTA counts -> offsets consumed by ti5_reset_observe for the ascending id list
  __shared__ int s_total;
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < nblk; ++i) {
      const int c = ((volatile int*)b.block_counts)[i];
      b.block_counts[i] = run;
      run += c;
    }
    b.block_counts[nblk] = run;
    s_total = run;
    g->n_reset = run;
    g->tickets[0] = 0;
    if (advance_force_flag && (p.flags & TI5_F_ADD_EXT_FORCE)) g->is_first_add_force = force_window ? 0 : 1;
  }
  __syncthreads();
  const int total = s_total;
  // extras["episode"] snapshot row of this step: new means if anything reset, else the previous row (A23)
  float* row = b.extras_log + (size_t)(step % TI5_LOG_ROWS) * TI5_LOG_COLS;
  const float* prev = b.extras_log + (size_t)((step + TI5_LOG_ROWS - 1) % TI5_LOG_ROWS) * TI5_LOG_COLS;
  __shared__ double s_track;
  if (threadIdx.x < TI5_NUM_TERMS) {
    const int t = threadIdx.x;
    float out = prev[t];
    if (total > 0) {
      // blocks without a reset never wrote their partial: skip them via their (now exclusive) offsets
      double acc = 0.0;
      for (int i = 0; i < nblk; ++i) {
        const int cnt = b.block_counts[i + 1] - b.block_counts[i];
        if (cnt > 0) acc += (double)((volatile float*)b.block_sums)[(size_t)i * TI5_LOG_COLS + t];
      }
      const float mean = (float)(acc / (double)total);
      out = sdiv(mean, p.max_episode_length_s, dm);
      if (t == T_TRACKING_LIN_VEL) s_track = acc / (double)total;
    }
    row[t] = out;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // lr:1160-1169 command curriculum, evaluated before the resets of this step (lr:537-538)
    if (total > 0 && (p.flags & TI5_F_COMMAND_CURRICULUM) && (counter % p.max_episode_length == 0)) {
      const float lhs = sdiv((float)s_track, (float)p.max_episode_length, dm);
      if (lhs > (float)(0.8 * p.tracking_lin_vel_scale)) {
        double lo = g->cmd_range[0][0] - 0.25, hi = g->cmd_range[0][1] + 0.5;
        const double lo_min = -p.cmd_curriculum_max / 2.0;
        g->cmd_range[0][0] = lo < lo_min ? lo_min : (lo > 0.0 ? 0.0 : lo);
        g->cmd_range[0][1] = hi < 0.0 ? 0.0 : (hi > p.cmd_curriculum_max ? p.cmd_curriculum_max : hi);
      }
    }
    row[LOG_MAX_COMMAND_X] = (float)g->cmd_range[0][1];
    row[LOG_N_RESET] = (float)total;
    if (total == 0) row[LOG_TERRAIN_LEVEL] = prev[LOG_TERRAIN_LEVEL];
  }
}

__global__ void __launch_bounds__(128)
post_physics_kernel(const __grid_constant__ Ti5Params p, const __grid_constant__ Ti5Buffers b,
                    const __grid_constant__ Ti5Rng r, int push_last) {
  const int N = p.num_envs;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = e < N;
  Ti5Globals* g = b.globals;
  const int64_t step = g->step_index;
  const int64_t counter = step + g->common_step_offset;   // common_step_counter after lr:471
  const bool first_force = g->is_first_add_force != 0;
  const bool philox = p.rng_mode == TI5_RNG_PHILOX;
  const int dm = p.div_mode;

  // window predicates are uniform over the grid (t1:193-215)
  bool push_window = false, force_window = false;
  if (p.flags & TI5_F_PUSH_ROBOTS) {
    int64_t i = counter / p.push_update_step;
    if (i >= p.n_push_dur) i = p.n_push_dur - 1;
    push_window = fmod((double)counter, (double)p.push_interval) <= p.push_duration[i];
  }
  if (p.flags & TI5_F_ADD_EXT_FORCE) {
    int64_t i = counter / p.add_update_step;
    if (i >= p.n_add_dur) i = p.n_add_dur - 1;
    force_window = fmod((double)counter, (double)p.ext_force_interval) <= p.add_duration[i];
  }

  bool reset = false;
  float esum[TI5_NUM_TERMS];   // episode sums of this env after this step (only used if it resets)
#pragma unroll
  for (int t = 0; t < TI5_NUM_TERMS; ++t) esum[t] = 0.0f;

  if (live) {
    // ---- simulator rows ---------------------------------------------------------------------
    float root[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) root[i] = b.root_states[(size_t)e * RB + i];
    float q[D], qd[D];
    {
      const float4* ds = reinterpret_cast<const float4*>(b.dof_state + (size_t)e * 2 * D);
#pragma unroll
      for (int i = 0; i < D / 2; ++i) {
        const float4 v = ds[i];
        q[2 * i] = v.x; qd[2 * i] = v.y; q[2 * i + 1] = v.z; qd[2 * i + 1] = v.w;
      }
    }
    FootState foot[2];
    float knee_xy[2][2];
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const float* rs = b.rigid_state + ((size_t)e * NB + p.feet[f]) * RB;
      foot[f].pos[0] = rs[0]; foot[f].pos[1] = rs[1]; foot[f].pos[2] = rs[2];
      foot[f].quat[0] = rs[3]; foot[f].quat[1] = rs[4]; foot[f].quat[2] = rs[5]; foot[f].quat[3] = rs[6];
      foot[f].wxy[0] = rs[10]; foot[f].wxy[1] = rs[11];
      const float* cf = b.contact_forces + ((size_t)e * NB + p.feet[f]) * 3;
      foot[f].force[0] = cf[0]; foot[f].force[1] = cf[1]; foot[f].force[2] = cf[2];
      const float* ks = b.rigid_state + ((size_t)e * NB + p.knees[f]) * RB;
      knee_xy[f][0] = ks[0]; knee_xy[f][1] = ks[1];
    }
    const float* tf = b.contact_forces + ((size_t)e * NB + p.term_body) * 3;
    const float term_force = sqrtf(tf[0] * tf[0] + tf[1] * tf[1] + tf[2] * tf[2]);
    const float* pf = b.contact_forces + ((size_t)e * NB + p.pen_body) * 3;
    const float pen_force = sqrtf(pf[0] * pf[0] + pf[1] * pf[1] + pf[2] * pf[2]);

    // ---- optional fused lag push of the last substep (lr:412-434) --------------------------
    if (push_last) {
      const int64_t j = (step - 1) * p.decimation + (p.decimation - 1);
      if (p.flags & TI5_F_ADD_DOF_LAG) {
        float4* row = reinterpret_cast<float4*>(b.dof_ring + ((size_t)ring_slot(j, p.dof_lag_len) * N + e) * (2 * D));
        row[0] = make_float4(q[0], q[1], q[2], q[3]);
        row[1] = make_float4(q[4], q[5], q[6], q[7]);
        row[2] = make_float4(q[8], q[9], q[10], q[11]);
        row[3] = make_float4(qd[0], qd[1], qd[2], qd[3]);
        row[4] = make_float4(qd[4], qd[5], qd[6], qd[7]);
        row[5] = make_float4(qd[8], qd[9], qd[10], qd[11]);
      }
    }

    // ---- lr:469-481 counters and derived base state ----------------------------------------
    const int64_t ep_len = b.episode_length_buf[e] + 1;
    b.episode_length_buf[e] = ep_len;
    const float bq[4] = {root[3], root[4], root[5], root[6]};
    const V3 lin = quat_rotate_inverse(bq, V3{root[7], root[8], root[9]});
    const V3 ang = quat_rotate_inverse(bq, V3{root[10], root[11], root[12]});
    const V3 grav = quat_rotate_inverse(bq, V3{0.0f, 0.0f, -1.0f});
    float eul[3];
    euler_xyz(bq, eul);
    float feul[2][3];
    euler_xyz(foot[0].quat, feul[0]);
    euler_xyz(foot[1].quat, feul[1]);
    foot[0].pitch = feul[0][1];
    foot[1].pitch = feul[1][1];
    reinterpret_cast<float4*>(b.base_quat)[e] = make_float4(bq[0], bq[1], bq[2], bq[3]);
    b.base_lin_vel[e * 3 + 0] = lin.x; b.base_lin_vel[e * 3 + 1] = lin.y; b.base_lin_vel[e * 3 + 2] = lin.z;
    b.base_ang_vel[e * 3 + 0] = ang.x; b.base_ang_vel[e * 3 + 1] = ang.y; b.base_ang_vel[e * 3 + 2] = ang.z;
    b.projected_gravity[e * 3 + 0] = grav.x; b.projected_gravity[e * 3 + 1] = grav.y; b.projected_gravity[e * 3 + 2] = grav.z;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      b.base_euler_xyz[e * 3 + i] = eul[i];
      b.feet_euler_xyz[e * 6 + i] = feul[0][i];
      b.feet_euler_xyz[e * 6 + 3 + i] = feul[1][i];
    }
    if (push_last && (p.flags & TI5_F_ADD_IMU_LAG)) {
      const int64_t j = (step - 1) * p.decimation + (p.decimation - 1);
      float* row = b.imu_ring + ((size_t)ring_slot(j, p.imu_lag_len) * N + e) * 6;
      row[0] = ang.x; row[1] = ang.y; row[2] = ang.z; row[3] = eul[0]; row[4] = eul[1]; row[5] = eul[2];
    }

    // ---- t1:183-184 phase counter and gait-schedule command resampling (pass 0) -------------
    int64_t phase_len = b.phase_length_buf[e] + 1;
    float4 cmd = reinterpret_cast<const float4*>(b.commands)[e];
    for (int gi = 0; gi < p.num_gaits; ++gi) {
      if (ep_len != (int64_t)b.gait_time[e * p.num_gaits + gi]) continue;
      const int kind = p.gait_kind[gi];
      float u[3];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        u[c] = philox ? philox_u(p.seed, (uint64_t)step, S_CMD + gi, e * 3 + c)
                      : r.cmd[((size_t)(0 * p.num_gaits + gi) * N + e) * 3 + c];
      const bool mx = kind == TI5_GAIT_WALK_SAGITTAL || kind == TI5_GAIT_WALK_OMNI;
      const bool my = kind == TI5_GAIT_WALK_LATERAL || kind == TI5_GAIT_WALK_OMNI;
      const bool mz = kind == TI5_GAIT_ROTATE || kind == TI5_GAIT_WALK_OMNI;
      cmd.x = mx ? affine((float)(g->cmd_range[0][1] - g->cmd_range[0][0]), (float)g->cmd_range[0][0], u[0]) : 0.0f;
      cmd.y = my ? affine((float)(g->cmd_range[1][1] - g->cmd_range[1][0]), (float)g->cmd_range[1][0], u[1]) : 0.0f;
      cmd.z = mz ? affine((float)(g->cmd_range[2][1] - g->cmd_range[2][0]), (float)g->cmd_range[2][0], u[2]) : 0.0f;
    }
    reinterpret_cast<float4*>(b.commands)[e] = cmd;
    const float cmd_norm = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y + cmd.z * cmd.z);
    const bool stand = cmd_norm <= p.stand_threshold;

    // ---- t1:193-203, 217-231 push window: overwrite the base velocity ----------------------
    if (p.flags & TI5_F_PUSH_ROBOTS) {
      float fx = 0.0f, fy = 0.0f, tq[3] = {0.0f, 0.0f, 0.0f};
      if (push_window) {
        float u[5];
#pragma unroll
        for (int c = 0; c < 5; ++c)
          u[c] = philox ? philox_u(p.seed, (uint64_t)step, S_PUSH, e * 5 + c) : r.push[(size_t)e * 5 + c];
        fx = affine(p.push_vel_w, p.push_vel_lo, u[0]);
        fy = affine(p.push_vel_w, p.push_vel_lo, u[1]);
#pragma unroll
        for (int c = 0; c < 3; ++c) tq[c] = affine(p.push_ang_w, p.push_ang_lo, u[2 + c]);
        root[7] = fx; root[8] = fy; root[10] = tq[0]; root[11] = tq[1]; root[12] = tq[2];
        float* rw = b.root_states + (size_t)e * RB;
        rw[7] = fx; rw[8] = fy; rw[10] = tq[0]; rw[11] = tq[1]; rw[12] = tq[2];
        b.rand_push_force[e * 3 + 0] = fx;
        b.rand_push_force[e * 3 + 1] = fy;
      } else {
        b.rand_push_force[e * 3 + 0] = 0.0f; b.rand_push_force[e * 3 + 1] = 0.0f; b.rand_push_force[e * 3 + 2] = 0.0f;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) b.rand_push_torque[e * 3 + c] = tq[c];
    }

    // ---- t1:205-215, 233-247 external force window ------------------------------------------
    if (p.flags & TI5_F_ADD_EXT_FORCE) {
      float af[3] = {0.0f, 0.0f, 0.0f}, at[3] = {0.0f, 0.0f, 0.0f};
      if (force_window) {
        if (first_force) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float uf = philox ? philox_u(p.seed, (uint64_t)step, S_EXT, e * 6 + c) : r.ext[(size_t)e * 6 + c];
            const float ut = philox ? philox_u(p.seed, (uint64_t)step, S_EXT, e * 6 + 3 + c) : r.ext[(size_t)e * 6 + 3 + c];
            b.ext_forces[e * 3 + c] = affine(p.ext_f_w[c], p.ext_f_lo[c], uf);
            b.ext_torques[e * 3 + c] = affine(p.ext_t_w, p.ext_t_lo, ut);
          }
        } else {
          const float s = stand ? 1.0f : 0.0f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            af[c] = b.ext_forces[e * 3 + c] * s;
            at[c] = b.ext_torques[e * 3 + c] * s;
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) { b.ext_forces[e * 3 + c] = 0.0f; b.ext_torques[e * 3 + c] = 0.0f; }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) { b.applied_force[e * 3 + c] = af[c]; b.applied_torque[e * 3 + c] = at[c]; }
    }

    // ---- lr:509-517 termination ---------------------------------------------------------------
    const bool time_out = ep_len > p.max_episode_length;
    reset = (term_force > 1.0f) || time_out;
    b.time_out_buf[e] = time_out ? 1 : 0;
    b.reset_buf[e] = reset ? 1 : 0;

    // ---- gait phase and stance mask (t1:80-107).  Side effect: standing envs restart the phase.
    if (stand) phase_len = 0;
    b.phase_length_buf[e] = phase_len;
    const float gait_start = b.gait_start[e];
    const float phase = (py_mod(sdiv((float)phase_len * p.dt, p.cycle_time, dm), 1.0f) + gait_start) * (stand ? 0.0f : 1.0f);
    const float sin_pos = sinf(TWO_PI_F * phase);
    float stance[2] = {sin_pos >= 0.0f ? 1.0f : 0.0f, sin_pos < 0.0f ? 1.0f : 0.0f};
    if (fabsf(sin_pos) < 0.1f) stance[0] = stance[1] = 1.0f;
    const bool contact[2] = {foot[0].force[2] > 5.0f, foot[1].force[2] > 5.0f};

    // ---- lr:654-680 reward sum, alphabetical term order ------------------------------------
    float rew = 0.0f;
    const uint32_t mask = p.term_mask;
    auto add_term = [&](int t, float v) {
      const float s = v * p.reward_scale[t];
      rew += s;
      const float acc = b.episode_sums[(size_t)t * N + e] + s;
      b.episode_sums[(size_t)t * N + e] = acc;
      esum[t] = acc;
      if (b.reward_terms) b.reward_terms[(size_t)t * N + e] = s;
    };

    float act[D];
    load12(b.actions, e, act);

    if (mask & (1u << T_ACTION_SMOOTHNESS)) {            // t1:877-892
      float la[D], lla[D];
      load12(b.last_actions, e, la);
      load12(b.last_last_actions, e, lla);
      float t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float d1 = (la[i] - act[i]) * 1.0f;
        const float d2 = ((act[i] + lla[i]) - 2.0f * la[i]) * 1.0f;
        t1 += d1 * d1;
        t2 += d2 * d2;
        t3 += fabsf(act[i] * 1.0f);
      }
      add_term(T_ACTION_SMOOTHNESS, (t1 + t2) + 0.05f * t3);
    }
    if (mask & (1u << T_BASE_ACC)) {                      // t1:717-724
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float d = b.last_root_vel[e * 6 + i] - root[7 + i];
        s += d * d;
      }
      add_term(T_BASE_ACC, expf(-sqrtf(s) * 3.0f));
    }
    if (mask & (1u << T_BASE_HEIGHT)) {                   // t1:706-715
      const float ground = (foot[0].pos[2] * stance[0] + foot[1].pos[2] * stance[1]) / (stance[0] + stance[1]);
      const float h = root[2] - (ground - 0.05f);
      add_term(T_BASE_HEIGHT, expf(-fabsf(h - p.base_height_target) * 100.0f));
    }
    if (mask & (1u << T_COLLISION)) {                     // t1:870-875
      add_term(T_COLLISION, 1.0f * (pen_force > 0.1f ? 1.0f : 0.0f));
    }
    float dq0[D];                                          // q - default (joint_diff)
#pragma unroll
    for (int i = 0; i < D; ++i) dq0[i] = q[i] - p.default_dof_pos[i];
    if (mask & (1u << T_DEFAULT_JOINT_POS)) {             // t1:686-703
      const float l = sqrtf((dq0[0] * dq0[0] + dq0[1] * dq0[1]) + dq0[5] * dq0[5]);
      const float rr = sqrtf((dq0[6] * dq0[6] + dq0[7] * dq0[7]) + dq0[11] * dq0[11]);
      const float yr = clampf((l + rr) - 0.1f, 0.0f, 50.0f);
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < D; ++i) s += dq0[i] * dq0[i];
      add_term(T_DEFAULT_JOINT_POS, expf(-yr * 100.0f) - 0.01f * sqrtf(s));
    }
    if (mask & ((1u << T_DOF_ACC))) {                      // t1:863-868
      float ldv[D];
      load12(b.last_dof_vel, e, ldv);
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float a = sdiv(ldv[i] - qd[i], p.dt, dm);
        s += a * a;
      }
      add_term(T_DOF_ACC, s);
    }
    if (mask & (1u << T_DOF_VEL)) {                       // t1:856-861
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < D; ++i) s += qd[i] * qd[i];
      add_term(T_DOF_VEL, s);
    }
    if (mask & (1u << T_FEET_AIR_TIME)) {                 // t1:642-657 (appendix A6, A8)
      const bool tiny = cmd_norm < 0.05f;
      float air_sum = 0.0f;
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        const float st = tiny ? 1.0f : stance[f];
        const bool filt = contact[f] || (st != 0.0f) || (b.last_contacts[e * 2 + f] != 0);
        b.contact_filt[e * 2 + f] = filt ? 1 : 0;
        b.last_contacts[e * 2 + f] = contact[f] ? 1 : 0;
        float air = b.feet_air_time[e * 2 + f];
        const float first = (air > 0.0f && filt) ? 1.0f : 0.0f;
        air += p.dt;
        air_sum += clampf(air, 0.0f, 0.5f) * first;
        b.feet_air_time[e * 2 + f] = air * (filt ? 0.0f : 1.0f);
      }
      add_term(T_FEET_AIR_TIME, air_sum);
    }
    if (mask & (1u << T_FEET_CLEARANCE)) {                // t1:793-814 (appendix A9)
      float s = 0.0f;
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        const float z = foot[f].pos[2];
        float h = b.feet_height[e * 2 + f] + (z - b.last_feet_z[e * 2 + f]);
        b.last_feet_z[e * 2 + f] = z;
        const float swing = 1.0f - stance[f];
        const float hit = (h > p.target_feet_height && h < p.target_feet_height_max) ? 1.0f : 0.0f;
        s += hit * swing;
        b.feet_height[e * 2 + f] = h * (contact[f] ? 0.0f : 1.0f);
      }
      add_term(T_FEET_CLEARANCE, s);
    }
    if (mask & (1u << T_FEET_CONTACT_FORCES)) {           // t1:679-684
      float s = 0.0f;
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        const float n = sqrtf((foot[f].force[0] * foot[f].force[0] + foot[f].force[1] * foot[f].force[1]) +
                              foot[f].force[2] * foot[f].force[2]);
        s += clampf(n - p.max_contact_force, 0.0f, 400.0f);
      }
      add_term(T_FEET_CONTACT_FORCES, s);
    }
    if (mask & (1u << T_FEET_CONTACT_NUMBER)) {           // t1:659-668 (appendix A14)
      float s = 0.0f;
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        const float st = stand ? 1.0f : stance[f];
        s += ((contact[f] ? 1.0f : 0.0f) == st) ? 1.0f : -0.3f;
      }
      add_term(T_FEET_CONTACT_NUMBER, s / 2.0f);
    }
    if (mask & (1u << T_FEET_DISTANCE)) {                 // t1:599-612
      add_term(T_FEET_DISTANCE, pair_distance_reward(foot[0].pos[0], foot[0].pos[1], foot[1].pos[0], foot[1].pos[1],
                                                     p.foot_min_dist, p.foot_max_dist));
    }
    if (mask & (1u << T_FEET_ROTATION)) {                 // t1:926-935 (appendix A11)
      const float rot = foot[0].pitch * foot[0].pitch + foot[1].pitch * foot[1].pitch;
      const float x = rot / 1.0f;
      add_term(T_FEET_ROTATION, 1.0f * expf(-(x * x)));
    }
    if (mask & (1u << T_FEET_STUMBLE)) {                  // t1:937-940
      bool any = false;
#pragma unroll
      for (int f = 0; f < 2; ++f)
        any = any || (sqrtf(foot[f].force[0] * foot[f].force[0] + foot[f].force[1] * foot[f].force[1]) >
                      5.0f * fabsf(foot[f].force[2]));
      add_term(T_FEET_STUMBLE, any ? 1.0f : 0.0f);
    }
    if (mask & (1u << T_FOOT_SLIP)) {                     // t1:630-640 (appendix A10)
      float s = 0.0f;
#pragma unroll
      for (int f = 0; f < 2; ++f)
        s += sqrtf(sqrtf(foot[f].wxy[0] * foot[f].wxy[0] + foot[f].wxy[1] * foot[f].wxy[1])) * (contact[f] ? 1.0f : 0.0f);
      add_term(T_FOOT_SLIP, s);
    }
    if (mask & (1u << T_JOINT_POS)) {                     // t1:576-596; ref_dof_pos of the PREVIOUS step (A3)
      float ref[D];
      load12(b.ref_dof_pos, e, ref);
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float d = q[i] - (stand ? p.default_dof_pos[i] : ref[i]);
        s += d * d;
      }
      const float n = sqrtf(s);
      const float v = expf(-2.0f * n) - 0.2f * clampf(n, 0.0f, 0.5f);
      add_term(T_JOINT_POS, stand ? 1.0f : v);
    }
    if (mask & (1u << T_KNEE_DISTANCE)) {                 // t1:615-628
      add_term(T_KNEE_DISTANCE, pair_distance_reward(knee_xy[0][0], knee_xy[0][1], knee_xy[1][0], knee_xy[1][1],
                                                     p.knee_min_dist, p.knee_max_dist));
    }
    if (mask & (1u << T_LOW_SPEED)) {                     // t1:816-847 (appendix A13)
      const float av = fabsf(lin.x), ac = fabsf(cmd.x);
      const bool slow = av < 0.5f * ac, fast = av > 1.2f * ac;
      float v = 0.0f;
      if (slow) v = -1.0f;
      if (fast) v = 0.0f;
      if (!(slow || fast)) v = 1.2f;
      if (signf(lin.x) != signf(cmd.x)) v = -2.0f;
      add_term(T_LOW_SPEED, v * (ac > 0.05f ? 1.0f : 0.0f));
    }
    if (mask & (1u << T_ORIENTATION)) {                   // t1:670-677
      const float a = expf(-(fabsf(eul[0]) + fabsf(eul[1])) * 10.0f);
      const float bb = expf(-sqrtf(grav.x * grav.x + grav.y * grav.y) * 20.0f);
      add_term(T_ORIENTATION, (a + bb) / 2.0f);
    }
    if (mask & (1u << T_STAND_STILL)) {                   // t1:899-915 (appendix A12)
      const int idx[8] = {0, 1, 2, 3, 5, 6, 7, 8};
      const float w[10] = {2.0f, 2.0f, 1.0f, 1.0f, 1.0f, 2.0f, 2.0f, 1.0f, 1.0f, 1.0f};
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float er = dq0[idx[i]] * w[i];
        s += er * er;
      }
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        const float er = foot[f].pitch * w[8 + f];
        s += er * er;
      }
      add_term(T_STAND_STILL, stand ? expf(-s) : 0.0f);
    }
    if (mask & (1u << T_STAND_SYSMETRY)) {                // t1:917-925
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float d = q[i] - q[5 + i];
        s += d * d;
      }
      add_term(T_STAND_SYSMETRY, stand ? expf(-s) : 0.0f);
    }
    if (mask & (1u << T_TORQUES)) {                       // t1:849-854
      float tau[D];
      load12(b.torques, e, tau);
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < D; ++i) s += tau[i] * tau[i];
      add_term(T_TORQUES, s);
    }
    const float ex = cmd.x - lin.x, ey = cmd.y - lin.y, ew = cmd.z - ang.z;
    if (mask & (1u << T_TRACK_VEL_HARD)) {                // t1:738-758
      const float le = sqrtf(ex * ex + ey * ey);
      const float ae = fabsf(ew);
      add_term(T_TRACK_VEL_HARD, (expf(-le * 10.0f) + expf(-ae * 10.0f)) / 2.0f - 0.2f * (le + ae));
    }
    if (mask & (1u << T_TRACKING_ANG_VEL)) {              // t1:776-790
      add_term(T_TRACKING_ANG_VEL, stand ? expf(-fabsf(ew) * p.tracking_sigma * 2.0f) : expf(-(ew * ew) * p.tracking_sigma));
    }
    if (mask & (1u << T_TRACKING_LIN_VEL)) {              // t1:760-774
      add_term(T_TRACKING_LIN_VEL, stand ? expf(-(fabsf(ex) + fabsf(ey)) * p.tracking_sigma * 2.0f)
                                         : expf(-(ex * ex + ey * ey) * p.tracking_sigma));
    }
    if (mask & (1u << T_VEL_MISMATCH_EXP)) {              // t1:726-736
      const float a = expf(-(lin.z * lin.z) * 10.0f);
      const float bb = expf(-sqrtf(ang.x * ang.x + ang.y * ang.y) * 5.0f);
      add_term(T_VEL_MISMATCH_EXP, (a + bb) / 2.0f);
    }
    if ((p.flags & TI5_F_ONLY_POSITIVE) && rew < 0.0f) rew = 0.0f;     // clip(min=0); NaN passes
    if (mask & (1u << T_TERMINATION)) {                   // lr:677-680, t1:894-896: added after the clip
      const float s = ((reset && !time_out) ? 1.0f : 0.0f) * p.reward_scale[T_TERMINATION];
      rew += s;
      const float acc = b.episode_sums[(size_t)T_TERMINATION * N + e] + s;
      b.episode_sums[(size_t)T_TERMINATION * N + e] = acc;
      esum[T_TERMINATION] = acc;
      if (b.reward_terms) b.reward_terms[(size_t)T_TERMINATION * N + e] = s;
    }
    b.rew_buf[e] = rew;
  }

  reset_bookkeeping(p, b, reset, esum, step, counter, force_window, true);
}

// Bookkeeping for an explicit `reset_idx(env_ids)` call (lr:450-455 `reset()`): the caller has
// written the mask into reset_buf; no physics, no rewards.
__global__ void __launch_bounds__(128)
reset_bookkeeping_kernel(const __grid_constant__ Ti5Params p, const __grid_constant__ Ti5Buffers b) {
  const int N = p.num_envs;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t step = b.globals->step_index;
  const int64_t counter = step + b.globals->common_step_offset;
  const bool reset = e < N && b.reset_buf[e] != 0;
  float esum[TI5_NUM_TERMS];
#pragma unroll
  for (int t = 0; t < TI5_NUM_TERMS; ++t) esum[t] = reset ? b.episode_sums[(size_t)t * N + e] : 0.0f;
  reset_bookkeeping(p, b, reset, esum, step, counter, false, false);
}

}  // namespace ti5

using namespace ti5;

extern "C" int ti5_reset_bookkeeping(const Ti5Params* p, const Ti5Buffers* b, void* stream) {
  TI5_CHECK_ARGS(p && b && p->num_envs > 0);
  TI5_CHECK_ARGS(p->env_block == 32 || p->env_block == 64 || p->env_block == 128);
  const int blocks = (p->num_envs + p->env_block - 1) / p->env_block;
  reset_bookkeeping_kernel<<<blocks, p->env_block, 0, (cudaStream_t)stream>>>(*p, *b);
  return ti5_check_launch("ti5_reset_bookkeeping");
}

extern "C" int ti5_post_physics(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, int push_last, void* stream) {
  TI5_CHECK_ARGS(p && b && p->num_envs > 0);
  TI5_CHECK_ARGS(p->env_block == 32 || p->env_block == 64 || p->env_block == 128);
  TI5_CHECK_ARGS(p->rng_mode == TI5_RNG_PHILOX || (r && r->cmd));
  TI5_CHECK_ARGS((p->term_mask & (1u << T_DOF_VEL_LIMITS)) == 0);   // the reference term reads a cfg field t1 lacks
  Ti5Rng rr = r ? *r : Ti5Rng{};
  const int blocks = (p->num_envs + p->env_block - 1) / p->env_block;
  post_physics_kernel<<<blocks, p->env_block, 0, (cudaStream_t)stream>>>(*p, *b, rr, push_last);
  return ti5_check_launch("ti5_post_physics");
}
