// Substep phase (10x per policy step): PD torques with action lag, and the DOF / IMU lag pushes.
//
// Replaces lr:393-394 (action clip), lr:1019-1074 (_compute_torques) and lr:412-434 (lag pushes).
// One thread per (env, 4 DOFs): every (N,12) array is read/written as coalesced float4, the AoS
// dof_state as two float4 per thread.  The reference shifts its (N,12,31)/(N,24,31)/(N,6,11)
// lag buffers by a full clone every substep (17.9 KB/env); here they are slot-major rings
// (len, N, width): a push is one coalesced row-block write, a lagged read a 48-byte gather.
#include <cstdlib>

#include "ti5_device.cuh"
#include "ti5_host.h"

namespace ti5 {

__global__ void __launch_bounds__(256) begin_step_kernel(const __grid_constant__ Ti5Params p, const __grid_constant__ Ti5Buffers b,
                                                        const float* __restrict__ actions_in) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0) b.globals->n_listed[(b.globals->step_index + 1) & 1] = 0;      // work list of the history clear, refilled by ti5_post_physics
  if (idx >= p.num_envs * D) return;
  b.actions[idx] = clampf(actions_in[idx], -p.clip_actions, p.clip_actions);
}

// One thread per (env, group of four DOFs): every per-DOF array moves as one float4 per thread, the
// interleaved dof_state as two, and one Philox call yields the four motor-strength multipliers.
__global__ void __launch_bounds__(256)
substep_kernel(const __grid_constant__ Ti5Params p, const __grid_constant__ Ti5Buffers b,
               const __grid_constant__ Ti5Rng r, const float* __restrict__ actions_in, int k_push, int k_torque, int phases) {
  chain_trigger();                                       // the next kernel of the step may become resident
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = p.num_envs;
  if (idx >= N * 3) return;
  const int e = idx / 3, gq = idx - e * 3, d0 = 4 * gq;
  const bool do_push = phases & TI5_SUB_PUSH, do_torque = phases & TI5_SUB_TORQUE;
  // chained launch: this grid may run ahead of the previous substep; `actions` (stored by the first substep) is
  // then read after chain_wait(), next to the lagged action row
  const bool late_actions = (phases & TI5_SUB_CHAINED) && actions_in == nullptr;
  const bool rg = p.flags & TI5_F_RAND_GAINS, fric = p.flags & TI5_F_RAND_COULOMB, rt = p.flags & TI5_F_RAND_TORQUE;
  const bool lagged = p.flags & TI5_F_ADD_LAG, imu = do_push && (p.flags & TI5_F_ADD_IMU_LAG);
  auto ld4 = [&](const float* base_ptr) { return reinterpret_cast<const float4*>(base_ptr)[idx]; };
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  // ---- every load that does not depend on another load, issued back to back ----------------------
  const int64_t step = b.globals->step_index + 1;       // index of the step in progress
  const float4* ds = reinterpret_cast<const float4*>(b.dof_state);
  const float4 s0 = ds[2 * idx], s1 = ds[2 * idx + 1];   // (q, qd) of DOFs d0..d0+3, interleaved
  // IMU push (lr:430-434): its three pieces — body-frame angular velocity + pitch, roll, yaw — are dealt out by
  // idx / N, so that a warp holds ONE kind of piece (no divergence between the atan2 / asin / rotation paths);
  // the env it works on, idx % N, is unrelated to the env of this thread's four DOFs.
  const int imu_piece = idx / N, imu_env = idx - imu_piece * N;
  float root[7] = {0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f};   // quat xyzw | ang vel
  if (imu) {
    const float* rp = b.root_states + (size_t)imu_env * RB;
#pragma unroll
    for (int i = 0; i < 4; ++i) root[i] = rp[3 + i];
    if (imu_piece == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) root[4 + i] = rp[10 + i];
    }
  }
  float4 a4 = zero4, kp4 = zero4, kd4 = zero4, off4 = zero4, vis4 = zero4, cou4 = zero4, u4 = zero4;
  int lag = 0;
  int64_t stamp = 0;
  if (do_torque) {
    if (actions_in != nullptr) {                         // fused lr:393-394 action clip (first substep only)
      a4 = ld4(actions_in);
      a4 = make_float4(clampf(a4.x, -p.clip_actions, p.clip_actions), clampf(a4.y, -p.clip_actions, p.clip_actions),
                       clampf(a4.z, -p.clip_actions, p.clip_actions), clampf(a4.w, -p.clip_actions, p.clip_actions));
      reinterpret_cast<float4*>(b.actions)[idx] = a4;
      if (idx == 0) b.globals->n_listed[step & 1] = 0;
    } else if (!late_actions) {
      a4 = ld4(b.actions);
    }
    off4 = ld4(b.motor_offsets);
    if (lagged) { lag = b.lag_timestep[e * 3 + 0]; stamp = b.ring_stamp[e]; }
    if (rg) { kp4 = ld4(b.p_gains_r); kd4 = ld4(b.d_gains_r); }
    if (fric) { vis4 = ld4(b.viscous); cou4 = ld4(b.coulomb); }
    if (rt && p.rng_mode != TI5_RNG_PHILOX) u4 = reinterpret_cast<const float4*>(r.torque)[(size_t)k_torque * N * 3 + idx];
  }
  const int64_t base = (step - 1) * p.decimation;        // pushes completed before this step
  const float q[4] = {s0.x, s0.z, s1.x, s1.z}, qd[4] = {s0.y, s0.w, s1.y, s1.w};

  // the lagged action row: pushed in an earlier step for most envs (lag > k) and cold in L2 by now — start fetching it
  float4* ring = reinterpret_cast<float4*>(b.act_ring);
  const int64_t jt = base + k_torque;                    // push index now
  int64_t jj = jt - lag;                                 // ... and the one the controller sees
  bool need_ring = do_torque && lagged && lag > 0;
  bool ring_live = need_ring && jj >= stamp && jj >= 0;  // rows pushed before the env's last reset read as zero (lr:606)
  const float4* ring_row = ring + (size_t)ring_slot(ring_live ? jj : 0, p.lag_len) * N * 3 + idx;
  // (with the per-substep re-draw of the lag, an option t1_cfg leaves off, the row is only known behind the wait)
  const bool relag = do_torque && lagged && (p.flags & TI5_F_LAG_PERSTEP);
  if (ring_live && !relag) prefetch_l2(ring_row);
  // the Philox draw and the IMU arithmetic need nothing from the previous kernel either
  if (do_torque && rt && p.rng_mode == TI5_RNG_PHILOX) u4 = philox_u4(p.seed, (uint64_t)step, S_TORQUE + k_torque, idx);
  float imu_val[4] = {0.f, 0.f, 0.f, 0.f};
  if (imu) {                                             // cat(base_ang_vel, base_euler_xyz), this thread's piece
    const float bq[4] = {root[0], root[1], root[2], root[3]};
    if (imu_piece == 0) {
      const V3 w = quat_rotate_inverse(bq, V3{root[4], root[5], root[6]});
      imu_val[0] = w.x; imu_val[1] = w.y; imu_val[2] = w.z;
      imu_val[3] = euler_pitch(bq);
    } else {
      // roll and yaw are the same atan2 form on different quaternion products (lr:30-33, 41-44)
      const float x = bq[0], y = bq[1], z = bq[2], w = bq[3];
      const bool yaw = imu_piece == 2;
      const float num = 2.0f * (yaw ? (w * z + x * y) : (w * x + y * z));
      const float den = yaw ? (((w * w + x * x) - y * y) - z * z) : (((w * w - x * x) - y * y) + z * z);
      imu_val[0] = wrap_angle(atan2f(num, den));
    }
  }

  // ---- everything below depends on (or must not overtake) the previous kernel of the step -------------
  chain_wait();
  if (relag) {
    // lr:1038-1043: a fresh lag index every substep, never more than one substep further back than the last one.  The
    // three threads of an env compute the same value from the copy of `last_lag` this substep reads (parity of the
    // push index); one of them stores the other copy for the next substep.
    const int lo = p.lag_range[0][0], hi = p.lag_range[0][1];
    int draw;
    if (p.rng_mode == TI5_RNG_PHILOX) {
      draw = lo + (int)(philox_u(p.seed, (uint64_t)step, S_LAGSTEP, e * 32 + k_torque) * (float)(hi - lo + 1));
      draw = draw > hi ? hi : draw;
    } else {
      draw = (int)r.lag_step[(size_t)k_torque * N + e];
    }
    const int last = b.last_lag[((size_t)(jt & 1) * N + e) * 5 + 0];
    lag = draw > last + 1 ? last + 1 : draw;
    if (gq == 0) {
      b.lag_timestep[e * 3 + 0] = lag;
      b.last_lag[((size_t)((jt + 1) & 1) * N + e) * 5 + 0] = lag;
    }
    jj = jt - lag;
    need_ring = lag > 0;
    ring_live = need_ring && jj >= stamp && jj >= 0;
    ring_row = ring + (size_t)ring_slot(ring_live ? jj : 0, p.lag_len) * N * 3 + idx;
  }
  // the one dependent load: the lagged action row (lr:1045)
  float4 t4 = zero4;
  if (ring_live) t4 = *ring_row;
  if (do_torque && late_actions) a4 = ld4(b.actions);

  if (do_push) {                                         // lr:412-434
    const int64_t j = base + k_push;
    if (p.flags & TI5_F_ADD_DOF_LAG) {
      float* row = b.dof_ring + ((size_t)ring_slot(j, p.dof_lag_len) * N + e) * (2 * D);
      *reinterpret_cast<float4*>(row + d0) = make_float4(q[0], q[1], q[2], q[3]);
      *reinterpret_cast<float4*>(row + D + d0) = make_float4(qd[0], qd[1], qd[2], qd[3]);
    }
    if (imu) {
      float* row = b.imu_ring + ((size_t)ring_slot(j, p.imu_lag_len) * N + imu_env) * 6;
      if (imu_piece == 0) {
        row[0] = imu_val[0]; row[1] = imu_val[1]; row[2] = imu_val[2];
        row[4] = imu_val[3];
      } else {
        row[imu_piece == 2 ? 5 : 3] = imu_val[0];
      }
    }
  }

  if (do_torque) {                                       // lr:1019-1074
    const float a[4] = {a4.x * p.action_scale, a4.y * p.action_scale, a4.z * p.action_scale, a4.w * p.action_scale};
    if (lagged) ring[(size_t)ring_slot(jt, p.lag_len) * N * 3 + idx] = make_float4(a[0], a[1], a[2], a[3]);
    const float target[4] = {need_ring ? t4.x : a[0], need_ring ? t4.y : a[1], need_ring ? t4.z : a[2], need_ring ? t4.w : a[3]};
    float kp[4] = {kp4.x, kp4.y, kp4.z, kp4.w}, kd[4] = {kd4.x, kd4.y, kd4.z, kd4.w};
    if (!rg) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { kp[i] = p.p_gains[d0 + i]; kd[i] = p.d_gains[d0 + i]; }
    }
    const float off[4] = {off4.x, off4.y, off4.z, off4.w};
    float tau[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float err = ((target[i] + p.default_dof_pos[d0 + i]) - q[i]) + off[i];
      tau[i] = kp[i] * err - kd[i] * qd[i];
    }
    if (fric) {
      const float vis[4] = {vis4.x, vis4.y, vis4.z, vis4.w}, cou[4] = {cou4.x, cou4.y, cou4.z, cou4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        tau[i] = tau[i] - vis[i] * qd[i];
        tau[i] = tau[i] - cou[i] * signf(qd[i]);
      }
    }
    if (rt) {
      const float u[4] = {u4.x, u4.y, u4.z, u4.w};
      float m[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        m[i] = affine(p.torque_multi_w, p.torque_multi_lo, u[i]);
        tau[i] = tau[i] * m[i];
      }
      reinterpret_cast<float4*>(b.torque_multi)[idx] = make_float4(m[0], m[1], m[2], m[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float lim = p.torque_limits[d0 + i];
      tau[i] = clampf(tau[i], -lim, lim);
    }
    reinterpret_cast<float4*>(b.torques)[idx] = make_float4(tau[0], tau[1], tau[2], tau[3]);
  }
}

}  // namespace ti5

using namespace ti5;

extern "C" int ti5_begin_step(const Ti5Params* p, const Ti5Buffers* b, const float* actions_in, void* stream) {
  TI5_CHECK_ARGS(p && b && actions_in && p->num_envs > 0);
  const int n = p->num_envs * D;
  begin_step_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p, *b, actions_in);
  return ti5_check_launch("ti5_begin_step");
}

static int launch_substep(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, const float* actions_in, int k, int phases,
                          void* stream) {
  TI5_CHECK_ARGS(p && b && p->num_envs > 0 && k >= 0 && k <= p->decimation && (phases & 3) != 0 && (phases & ~7) == 0);
  TI5_CHECK_ARGS(!(phases & TI5_SUB_TORQUE) || k < p->decimation);
  TI5_CHECK_ARGS(p->decimation <= 16);       // Philox sites S_TORQUE + k must stay below S_CMD
  TI5_CHECK_ARGS(!(p->flags & TI5_F_LAG_PERSTEP) || (b->last_lag && (p->rng_mode == TI5_RNG_PHILOX || (r && r->lag_step))));
  TI5_CHECK_ARGS(!(phases & TI5_SUB_TORQUE) || p->rng_mode == TI5_RNG_PHILOX || !(p->flags & TI5_F_RAND_TORQUE) ||
                 (r && r->torque));
  Ti5Rng rr = r ? *r : Ti5Rng{};
  const int n = p->num_envs * 3;
  // fused form: push the result of simulator substep k-1, then the torque of substep k
  const int k_push = (phases & TI5_SUB_TORQUE) ? k - 1 : k;
  TI5_CHECK_ARGS(!(phases & TI5_SUB_PUSH) || k_push >= 0);
  static const int block = getenv("TI5_SUBSTEP_BLOCK") ? atoi(getenv("TI5_SUBSTEP_BLOCK")) : 64;
  ti5_set_carveout(substep_kernel, ti5_small_grid(p));
  (void)ti5_launch(substep_kernel, dim3((n + block - 1) / block), dim3(block), 0, stream, (phases & TI5_SUB_CHAINED) != 0, *p, *b, rr,
                   actions_in, k_push, k, phases);
  return ti5_check_launch("ti5_substep");
}

extern "C" int ti5_substep(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, int k, int phases, void* stream) {
  return launch_substep(p, b, r, nullptr, k, phases, stream);
}

extern "C" int ti5_first_substep(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, const float* actions_in, void* stream) {
  TI5_CHECK_ARGS(actions_in != nullptr);
  return launch_substep(p, b, r, actions_in, 0, TI5_SUB_TORQUE, stream);
}

extern "C" int ti5_torque_substep(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, int k, void* stream) {
  return ti5_substep(p, b, r, k, TI5_SUB_TORQUE, stream);
}

extern "C" int ti5_lag_push(const Ti5Params* p, const Ti5Buffers* b, int k, void* stream) {
  return ti5_substep(p, b, nullptr, k, TI5_SUB_PUSH, stream);
}
