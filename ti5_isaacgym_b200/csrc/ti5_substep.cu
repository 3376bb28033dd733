// Substep phase (10x per policy step): PD torques with action lag, and the DOF / IMU lag pushes.
//
// Replaces lr:393-394 (action clip), lr:1019-1074 (_compute_torques) and lr:412-434 (lag pushes).
// One thread per (env, DOF) element: every (N,12) array is read/written fully coalesced, the AoS
// dof_state is read as one float2 per thread.  The reference shifts its (N,12,31)/(N,24,31)/(N,6,11)
// lag buffers by a full clone every substep (17.9 KB/env); here they are slot-major rings
// (len, N, width): a push is one coalesced row-block write, a lagged read a 48-byte gather.
#include "ti5_device.cuh"
#include "ti5_host.h"

namespace ti5 {

__global__ void __launch_bounds__(256) begin_step_kernel(const __grid_constant__ Ti5Params p, const __grid_constant__ Ti5Buffers b,
                                                        const float* __restrict__ actions_in) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0) b.globals->step_index += 1;   // nobody in this launch reads it
  if (idx >= p.num_envs * D) return;
  b.actions[idx] = clampf(actions_in[idx], -p.clip_actions, p.clip_actions);
}

// value pushed to IMU-lag column c of env e: cat(base_ang_vel, base_euler_xyz) (lr:430-434)
__device__ __forceinline__ float imu_component(const float* __restrict__ root, int c) {
  float q[4] = {root[3], root[4], root[5], root[6]};
  if (c < 3) {
    const V3 w = quat_rotate_inverse(q, V3{root[10], root[11], root[12]});
    return c == 0 ? w.x : (c == 1 ? w.y : w.z);
  }
  return c == 3 ? euler_roll(q) : (c == 4 ? euler_pitch(q) : euler_yaw(q));
}

__global__ void __launch_bounds__(256)
substep_kernel(const __grid_constant__ Ti5Params p, const __grid_constant__ Ti5Buffers b,
               const __grid_constant__ Ti5Rng r, int k_push, int k_torque, int phases) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = p.num_envs;
  if (idx >= N * D) return;
  const int e = idx / D, d = idx - e * D;
  const int64_t step = b.globals->step_index;           // >= 1 inside a step
  const int64_t base = (step - 1) * p.decimation;        // pushes completed before this step
  const float2 qs = reinterpret_cast<const float2*>(b.dof_state)[idx];   // (pos, vel)

  if (phases & TI5_SUB_PUSH) {
    const int64_t j = base + k_push;
    if (p.flags & TI5_F_ADD_DOF_LAG) {
      float* row = b.dof_ring + ((size_t)ring_slot(j, p.dof_lag_len) * N + e) * (2 * D);
      row[d] = qs.x;
      row[D + d] = qs.y;
    }
    if ((p.flags & TI5_F_ADD_IMU_LAG) && d < 6) {
      b.imu_ring[((size_t)ring_slot(j, p.imu_lag_len) * N + e) * 6 + d] = imu_component(b.root_states + (size_t)e * RB, d);
    }
  }

  if (phases & TI5_SUB_TORQUE) {
    const int64_t j = base + k_torque;
    const float a = b.actions[idx] * p.action_scale;
    float target = a;
    if (p.flags & TI5_F_ADD_LAG) {
      b.act_ring[((size_t)ring_slot(j, p.lag_len) * N) * D + idx] = a;
      const int lag = b.lag_timestep[e * 3 + 0];
      if (lag > 0) {
        const int64_t jj = j - lag;     // push index the controller sees; rows pushed before the
        target = (jj >= b.ring_stamp[e] && jj >= 0)   // env's last reset read as zero (lr:606)
                     ? b.act_ring[((size_t)ring_slot(jj, p.lag_len) * N) * D + idx] : 0.0f;
      }
    }
    const bool rg = p.flags & TI5_F_RAND_GAINS;
    const float kp = rg ? b.p_gains_r[idx] : p.p_gains[d];
    const float kd = rg ? b.d_gains_r[idx] : p.d_gains[d];
    const float err = ((target + p.default_dof_pos[d]) - qs.x) + b.motor_offsets[idx];
    float tau = kp * err - kd * qs.y;
    if (p.flags & TI5_F_RAND_COULOMB) {
      tau = tau - b.viscous[idx] * qs.y;
      tau = tau - b.coulomb[idx] * signf(qs.y);
    }
    if (p.flags & TI5_F_RAND_TORQUE) {
      const float u = p.rng_mode == TI5_RNG_PHILOX ? philox_u(p.seed, (uint64_t)step, S_TORQUE + k_torque, idx)
                                                   : r.torque[(size_t)k_torque * N * D + idx];
      const float m = affine(p.torque_multi_w, p.torque_multi_lo, u);
      b.torque_multi[idx] = m;
      tau = tau * m;
    }
    const float lim = p.torque_limits[d];
    b.torques[idx] = clampf(tau, -lim, lim);
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

extern "C" int ti5_substep(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, int k, int phases, void* stream) {
  TI5_CHECK_ARGS(p && b && p->num_envs > 0 && k >= 0 && k <= p->decimation && (phases & 3) != 0);
  TI5_CHECK_ARGS(!(phases & TI5_SUB_TORQUE) || k < p->decimation);
  TI5_CHECK_ARGS(!(phases & TI5_SUB_TORQUE) || p->rng_mode == TI5_RNG_PHILOX || !(p->flags & TI5_F_RAND_TORQUE) ||
                 (r && r->torque));
  Ti5Rng rr = r ? *r : Ti5Rng{};
  const int n = p->num_envs * D;
  // fused form: push the result of simulator substep k-1, then the torque of substep k
  const int k_push = (phases & TI5_SUB_TORQUE) ? k - 1 : k;
  TI5_CHECK_ARGS(!(phases & TI5_SUB_PUSH) || k_push >= 0);
  substep_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p, *b, rr, k_push, k, phases);
  return ti5_check_launch("ti5_substep");
}

extern "C" int ti5_torque_substep(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, int k, void* stream) {
  return ti5_substep(p, b, r, k, TI5_SUB_TORQUE, stream);
}

extern "C" int ti5_lag_push(const Ti5Params* p, const Ti5Buffers* b, int k, void* stream) {
  return ti5_substep(p, b, nullptr, k, TI5_SUB_PUSH, stream);
}
