// Post-physics phase, part 1 (once per policy step): counters, derived base state, command
// schedule, push / external-force windows, termination, the reward sum, and the reset
// bookkeeping.  One thread per env; the AoS simulator rows are read once into registers and
// every (TERMS, N) / (N, k) product array is written coalesced.
//
// Replaces, in the reference's order:  lr:464-481 (post_physics_step head), t1:179-215
// (_post_physics_step_callback), lr:509-517 (check_termination), lr:654-680 + t1:572-946
// (compute_reward and the reward terms), and the reductions reset_idx needs before it can run
// (lr:490 count, t1:530-541 episode means, lr:1160-1169 command curriculum).
#include <cstddef>

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
static __device__ __noinline__ float pair_distance_reward(float ax, float ay, float bx, float by, float lo, float hi) {
  const float dx = ax - bx, dy = ay - by;
  const float d = sqrtf(dx * dx + dy * dy);
  const float near_ = clampf(d - lo, -0.5f, 0.0f);
  const float far_ = clampf(d - hi, 0.0f, 0.5f);
  return (expf_call(-fabsf(near_) * 100.0f) + expf_call(-fabsf(far_) * 100.0f)) / 2.0f;
}

// Reset bookkeeping shared by ti5_post_physics and ti5_reset_bookkeeping: the CTA's count of flagged envs,
// the arrival-order work list for the history clear, and — only if the CTA has a flagged env — the partial
// sums of their episode sums (t1:531-533).  Nothing here waits on another CTA: ti5_reset_observe, which runs
// after this grid has completed, turns the per-CTA counts into offsets / totals itself.
// `sums` is the CTA's shared tile [term][tb] (post-physics) or null (read episode_sums from memory).
// `listed`: the caller has already put its flagged envs on the work list (post_physics_kernel does so the moment the
// flag is known: the returning atomic is a ~1 us round trip, and the CTAs that hold a flagged env finish last).
static __device__ __noinline__ void reset_bookkeeping(const Ti5Params& p, const Ti5Buffers& b, bool reset, int e, int le,
                                                      const float* sums, int tb, int64_t step, bool listed) {
  __shared__ float s_red[32][TI5_NUM_TERMS];
  const int N = p.num_envs, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  if (reset && !listed) b.reset_list[atomicAdd(&b.globals->n_listed[step & 1], 1)] = e;
  const int total = __syncthreads_count(reset);          // one barrier: only the count is needed here
  if (tid == 0) b.block_counts[blockIdx.x] = total;
  if (total == 0) return;
  // lanes over terms: lane t adds up term t of the warp's flagged envs (a handful), then the warps are folded
  const unsigned flagged = __ballot_sync(0xffffffffu, reset);
  float acc = 0.0f;
  const int warp_le0 = le - lane;                      // first tile-env of this warp (role-0 warps hold the flags)
  for (unsigned m = flagged; m; m &= m - 1) {
    const int src = __ffs(m) - 1;
    if (lane < TI5_NUM_TERMS && (p.term_mask & (1u << lane)))
      acc += sums ? sums[lane * tb + warp_le0 + src] : b.episode_sums[(size_t)lane * N + (e - lane + src)];
  }
  if (lane < TI5_NUM_TERMS) s_red[warp][lane] = acc;
  __syncthreads();
  if (tid < TI5_NUM_TERMS) {
    float v = 0.0f;
    for (int w = 0; w < (nt >> 5); ++w) v += s_red[w][tid];
    b.block_sums[(size_t)blockIdx.x * TI5_LOG_COLS + tid] = v;
  }
}

// Shared-memory tile of the per-env inputs of one CTA (TB = env_block consecutive envs), filled by the TMA
// engine.  The 20 per-env arrays are described by a table (source pointer, bytes per env, byte offset per env
// inside the tile) built on the host, so thread k issues bulk copy k without any per-array code; the tile also
// holds one column per reward term of the episode sums, the unscaled term values and the mbarrier.
constexpr int POST_CHUNKS = 20;
enum PostChunk { C_ROOT = 0, C_DOF, C_CONTACT, C_RIGID, C_ACT, C_LAST_ACT, C_LAST_LAST_ACT, C_LAST_DOF_VEL, C_TORQUES, C_REF,
                 C_LAST_ROOT_VEL, C_CMD, C_GAIT_START, C_AIR, C_FEET_H, C_LAST_Z, C_EP_LEN, C_PHASE_LEN, C_GAIT_TIME,
                 C_LAST_CONTACTS };
constexpr int STAGED_BODIES = 4;     // contact tile rows: [termination, penalised, foot 0, foot 1]; rigid tile rows: [foot 0, foot 1, knee 0, knee 1]
constexpr int CON_STRIDE = STAGED_BODIES * 3 + 1, RIG_STRIDE = STAGED_BODIES * RB + 1;   // floats per env in the tile: odd, conflict-free by env
struct PostSrc {
  const void* ptr[POST_CHUNKS];
  int32_t rowb[POST_CHUNKS];          // bytes per env
  int32_t off[POST_CHUNKS + 1];       // byte offset per env of chunk k inside the tile (prefix sum of rowb)
};

static PostSrc make_post_src(const Ti5Params& p, const Ti5Buffers& b) {
  PostSrc s;
  const void* ptr[POST_CHUNKS] = {b.root_states, b.dof_state, b.contact_forces, b.rigid_state, b.actions, b.last_actions,
                                  b.last_last_actions, b.last_dof_vel, b.torques, b.ref_dof_pos, b.last_root_vel, b.commands,
                                  b.gait_start, b.feet_air_time, b.feet_height, b.last_feet_z, b.episode_length_buf,
                                  b.phase_length_buf, b.gait_time, b.last_contacts};
  // C_CONTACT and C_RIGID are NOT copied whole: of the 13 bodies the phase reads the contact force of the termination /
  // penalised body and the two feet, and the state of the two feet and the two knees — STAGED_BODIES rows each, gathered
  // with 4-byte asynchronous copies (stage_body_rows)
  const int rowb[POST_CHUNKS] = {RB * 4, 2 * D * 4, CON_STRIDE * 4, RIG_STRIDE * 4, D * 4, D * 4, D * 4, D * 4, D * 4, D * 4,
                                 6 * 4, 4 * 4, 4, 2 * 4, 2 * 4, 2 * 4, 8, 8, p.num_gaits * 4, 2};
  int o = 0;
  for (int k = 0; k < POST_CHUNKS; ++k) { s.ptr[k] = ptr[k]; s.rowb[k] = rowb[k]; s.off[k] = o; o += rowb[k]; }
  s.off[POST_CHUNKS] = o;
  return s;
}

struct PostTile {
  unsigned char* base;
  float *sums, *vals;
  uint64_t* bar;
  int tb;
  const PostSrc* src;
  template <class T> __device__ __forceinline__ T* at(int chunk) const { return reinterpret_cast<T*>(base + (size_t)src->off[chunk] * tb); }
};

constexpr int FOOT_PARTS = 9;        // FootPart entries handed from one foot role to the other (post_physics_kernel)
// The rows of the AoS contact-force (N,13,3) and rigid-body (N,13,13) tensors the phase uses, into the tile: 4-byte
// asynchronous copies (a 12- or 52-byte row at a 4-byte-aligned address is nothing the bulk engine takes), issued by
// `nthreads` threads together with the bulk copies of the other arrays; 244 of 832 bytes per env are fetched.
__device__ __forceinline__ void stage_body_rows(const Ti5Params& p, const Ti5Buffers& b, float* t_contact, float* t_rigid,
                                                int e0, int n_tile, int t, int nthreads) {
  // twelve work items per env — four 3-float contact rows, four 13-float rigid-body rows as halves of 7 + 6 floats — so
  // that a thread forms one row address and then issues its copies at immediate offsets (the first form, one item per
  // FLOAT with two divisions each, was 10 % of the kernel's executed instructions, in front of the tile wait)
  constexpr int ITEMS = 3 * STAGED_BODIES;
#pragma unroll 1
  for (int it = t; it < n_tile * ITEMS; it += nthreads) {
    const int en = it / ITEMS, r = it - en * ITEMS;
    const float* src;
    float* dst;
    int n;
    if (r < STAGED_BODIES) {
      const int body = r == 0 ? p.term_body : r == 1 ? p.pen_body : p.feet[r - 2];
      src = b.contact_forces + ((size_t)(e0 + en) * NB + body) * 3;
      dst = t_contact + en * CON_STRIDE + r * 3;
      n = 3;
    } else {
      const int rr = (r - STAGED_BODIES) >> 1, half = (r - STAGED_BODIES) & 1;
      const int body = rr < 2 ? p.feet[rr] : p.knees[rr - 2];
      src = b.rigid_state + ((size_t)(e0 + en) * NB + body) * RB + half * 7;
      dst = t_rigid + en * RIG_STRIDE + rr * RB + half * 7;
      n = half ? RB - 7 : 7;
    }
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (j < n) cp_async4(dst + j, src + j);
  }
}

// `fused_dec` > 0 (ti5_fused_step): behind the tile, per (substep, env, four DOFs) item one float4 of torque-multiplier
// uniforms (drawn by the role threads while the tile is in flight) and one float4 for the lagged action row a substep
// reads (fetched by 16-byte asynchronous copies before the substep loop starts)
__host__ __device__ inline size_t post_tile_bytes(int tb, int per_env_bytes, int fused_dec = 0) {
  return (size_t)tb * per_env_bytes + 2 * (size_t)TI5_NUM_TERMS * tb * 4 + 16 + (size_t)FOOT_PARTS * tb * 4 +
         2 * (size_t)fused_dec * 3 * tb * 16;
}

// lr:393-434 when no simulator runs between the substeps, for one (env, group of four DOFs): the action clip, then
// [_compute_torques, DOF-lag push] for the substeps k = kh, kh + WORKER_SPLIT, ... with the joint state, gains and offsets
// read once.  Same arithmetic and the same Philox counters as DEC launches of substep_kernel (ti5_substep.cu); every
// substep's torques are stored (torques_substeps[k]), the last substep's also where the unfused kernels leave them.
// The IMU-lag pushes, whose values equal the derived base state, are left to role 0.  The loop is rolled on purpose:
// a warp runs it once, so every instruction of an unrolled body would be a cold instruction-cache line.
// WS = threads per (env, four DOFs): the substeps are dealt out round-robin.  2 on small grids (the kernel is bound by
// the length of a thread's chain there), 1 on large ones (by the number of instructions: half as many worker set-ups,
// and a CTA small enough for three per SM)
// named barriers (0 is __syncthreads)
constexpr int BAR_TORQUES = 1;      // substep workers arrive, R_JOINT_B waits
constexpr int BAR_ROLES = 2;        // the role threads among themselves while the workers run (FUSED)
constexpr int BAR_FEET = 3;         // R_FOOT1 arrives, R_FOOT0 waits
constexpr int BAR_DRAWS = 4;        // the role threads arrive (torque-multiplier uniforms staged), the substep workers wait (FUSED)
__device__ __forceinline__ void bar_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// Two halves around the CTA's BAR_DRAWS barrier: setup() issues every load of the item (joint state, actuator arrays,
// and — by 16-byte asynchronous copies into `s_old` — all lagged action rows this thread will read); run() is the
// substep loop, which takes the torque-multiplier uniforms of (substep, item) from `s_u4`, where the role threads
// staged them (same Philox counters / pool entries) while the tile was in flight.
template <int WORKER_SPLIT>
struct SubstepWorker {
  float4 q4, qd4, as4, kp4, kd4, off4, vis4, cou4;
  int lag, ws, dsl;

  __device__ __forceinline__ void setup(const Ti5Params& p, const Ti5Buffers& b, const float* __restrict__ actions_in, int64_t step,
                                        int idx, int e, int gq, int kh, float* tile_act, float4* s_old, int item, int items) {
    const int N = p.num_envs, d0 = 4 * gq, dec = p.decimation;
    const bool rg = p.flags & TI5_F_RAND_GAINS, fric = p.flags & TI5_F_RAND_COULOMB, lagged = p.flags & TI5_F_ADD_LAG;
    auto ld4 = [&](const float* base_ptr) { return reinterpret_cast<const float4*>(base_ptr)[idx]; };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // ---- loads, all independent ----------------------------------------------------------------------
    const float4* ds = reinterpret_cast<const float4*>(b.dof_state);
    const float4 s0 = ds[2 * idx], s1 = ds[2 * idx + 1];
    float4 a4 = ld4(actions_in);
    off4 = ld4(b.motor_offsets);
    kp4 = kd4 = vis4 = cou4 = zero4;
    lag = 0;
    int64_t stamp = 0;
    if (lagged) { lag = b.lag_timestep[e * 3 + 0]; stamp = b.ring_stamp[e]; }
    if (rg) { kp4 = ld4(b.p_gains_r); kd4 = ld4(b.d_gains_r); }
    else { kp4 = make_float4(p.p_gains[d0], p.p_gains[d0 + 1], p.p_gains[d0 + 2], p.p_gains[d0 + 3]);
           kd4 = make_float4(p.d_gains[d0], p.d_gains[d0 + 1], p.d_gains[d0 + 2], p.d_gains[d0 + 3]); }
    if (fric) { vis4 = ld4(b.viscous); cou4 = ld4(b.coulomb); }
    const int64_t base = (step - 1) * dec;                   // pushes completed before this step
    const float4* ring = reinterpret_cast<const float4*>(b.act_ring);
    const size_t ring_row = (size_t)N * 3;
    // ring slots: one remainder per ring, then increments
    const int alen = p.lag_len, dlen = p.dof_lag_len;
    ws = (int)fast_mod(base + kh, alen);                     // slot the action of substep k is pushed to
    dsl = (int)fast_mod(base + kh, dlen);                    // slot of the DOF-lag push after substep k
    // the lagged action row of a substep that looks back past the start of this step (lr:1045); rows pushed before the
    // env's last reset read as zero (lr:606).  All of this thread's old rows (pushed up to three steps ago, evicted
    // from the L2 since) are fetched up front; every copy of every worker of the CTA is complete before the barrier,
    // i.e. before any push of this step, so no slot is overwritten in front of its reader.
    if (lagged) {
      int slot = ws - lag % alen;                            // slot of the lagged row substep kh looks at
      if (slot < 0) slot += alen;
      // row j = base + k - lag is live iff j >= max(stamp, 0), i.e. k >= kmin
      const int64_t first_live = (stamp > 0 ? stamp : 0) - (base - lag);
      const int kmin = first_live <= 0 ? 0 : (first_live >= dec ? dec : (int)first_live);
#pragma unroll 1
      for (int k = kh; k < dec && k < lag; k += WORKER_SPLIT) {
        float4* dst = s_old + k * items + item;
        if (k >= kmin) cp_async16(dst, ring + (size_t)slot * ring_row + idx);
        else *dst = zero4;
        slot += WORKER_SPLIT;
        while (slot >= alen) slot -= alen;
      }
    }
    a4 = make_float4(clampf(a4.x, -p.clip_actions, p.clip_actions), clampf(a4.y, -p.clip_actions, p.clip_actions),
                     clampf(a4.z, -p.clip_actions, p.clip_actions), clampf(a4.w, -p.clip_actions, p.clip_actions));   // lr:393-394
    if (kh == 0) {
      reinterpret_cast<float4*>(b.actions)[idx] = a4;
      *reinterpret_cast<float4*>(tile_act + d0) = a4;
    }
    q4 = make_float4(s0.x, s0.z, s1.x, s1.z);
    qd4 = make_float4(s0.y, s0.w, s1.y, s1.w);
    as4 = make_float4(a4.x * p.action_scale, a4.y * p.action_scale, a4.z * p.action_scale, a4.w * p.action_scale);
    cp_async_wait_all();                                     // this thread's old rows are in shared memory
  }

  __device__ __forceinline__ void run(const Ti5Params& p, const Ti5Buffers& b, const Ti5Rng& r, int64_t step, int idx, int e,
                                      int gq, int kh, float* tile_tau, const float4* s_u4, const float4* s_old, int item,
                                      int items) {
    const int N = p.num_envs, d0 = 4 * gq, dec = p.decimation;
    const bool fric = p.flags & TI5_F_RAND_COULOMB, rt = p.flags & TI5_F_RAND_TORQUE, lagged = p.flags & TI5_F_ADD_LAG;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* ring = reinterpret_cast<float4*>(b.act_ring);
    const size_t ring_row = (size_t)N * 3;
    const int alen = p.lag_len, dlen = p.dof_lag_len;
    auto adv = [](int s, int len) {          // (s + WORKER_SPLIT) mod len without a division (len >= 1)
      s += WORKER_SPLIT;
      while (s >= len) s -= len;
      return s;
    };
    const float q[4] = {q4.x, q4.y, q4.z, q4.w}, qd[4] = {qd4.x, qd4.y, qd4.z, qd4.w};
    const float a[4] = {as4.x, as4.y, as4.z, as4.w};
    const float kp[4] = {kp4.x, kp4.y, kp4.z, kp4.w}, kd[4] = {kd4.x, kd4.y, kd4.z, kd4.w};
    const float off[4] = {off4.x, off4.y, off4.z, off4.w};
    const float vis[4] = {vis4.x, vis4.y, vis4.z, vis4.w}, cou[4] = {cou4.x, cou4.y, cou4.z, cou4.w};
    const bool from_ring = lagged && lag > 0;
#pragma unroll 1
    for (int k = kh; k < dec; k += WORKER_SPLIT) {
      const float4 t4 = lag <= k ? as4 : s_old[k * items + item];      // a row this step pushed itself: from registers
      // lr:1019-1074 torque of substep k
      const float4 u4 = rt ? s_u4[k * items + item] : zero4;
      if (lagged) ring[(size_t)ws * ring_row + idx] = as4;
      const float target[4] = {from_ring ? t4.x : a[0], from_ring ? t4.y : a[1], from_ring ? t4.z : a[2], from_ring ? t4.w : a[3]};
      float tau[4], m[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float err = ((target[i] + p.default_dof_pos[d0 + i]) - q[i]) + off[i];
        tau[i] = kp[i] * err - kd[i] * qd[i];
      }
      if (fric) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          tau[i] = tau[i] - vis[i] * qd[i];
          tau[i] = tau[i] - cou[i] * signf(qd[i]);
        }
      }
      if (rt) {
        const float u[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          m[i] = affine(p.torque_multi_w, p.torque_multi_lo, u[i]);
          tau[i] = tau[i] * m[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float lim = p.torque_limits[d0 + i];
        tau[i] = clampf(tau[i], -lim, lim);
      }
      const float4 tau4 = make_float4(tau[0], tau[1], tau[2], tau[3]);
      if (b.torques_substeps) reinterpret_cast<float4*>(b.torques_substeps)[(size_t)k * ring_row + idx] = tau4;
      if (k == dec - 1) {                                    // what the unfused sequence leaves behind (lr:401, 1072)
        reinterpret_cast<float4*>(b.torques)[idx] = tau4;
        if (rt) reinterpret_cast<float4*>(b.torque_multi)[idx] = make_float4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<float4*>(tile_tau + d0) = tau4;
      }
      // lr:412-418 DOF-lag push after (what would be) simulator substep k
      if (p.flags & TI5_F_ADD_DOF_LAG) {
        float* row = b.dof_ring + ((size_t)dsl * N + e) * (2 * D);
        *reinterpret_cast<float4*>(row + d0) = q4;
        *reinterpret_cast<float4*>(row + D + d0) = qd4;
      }
      ws = adv(ws, alen);
      dsl = adv(dsl, dlen);
    }
  }
};

// The torque-multiplier uniforms of all (substep, item) pairs of the CTA (lr:1065-1068; one float4 = the four DOFs of
// an item), staged by `nthreads` role threads while the tile is in flight: NCH Philox chains per thread side by side
// (a single chain is a ~500-cycle dependent sequence; in the substep loop it was the largest part of a worker's time).
__device__ __forceinline__ void stage_torque_uniforms(const Ti5Params& p, const Ti5Rng& r, int64_t step, int e0, int items,
                                                      float4* s_u4, int t, int nthreads) {
  constexpr int NCH = 5;
  const int total = p.decimation * items;
  if (p.rng_mode != TI5_RNG_PHILOX) {
    const size_t ring_row = (size_t)p.num_envs * 3;
    for (int j = t; j < total; j += nthreads) {
      const int k = j / items, item = j - k * items;
      const size_t idx = (size_t)e0 * 3 + item;
      s_u4[j] = idx < ring_row ? reinterpret_cast<const float4*>(r.torque)[(size_t)k * ring_row + idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  const uint32_t key0 = (uint32_t)p.seed, key1 = (uint32_t)(p.seed >> 32);
#pragma unroll 1
  for (int j0 = t; j0 < total; j0 += NCH * nthreads) {
    uint32_t c0[NCH], c1[NCH], c2[NCH], c3[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int j = j0 + c * nthreads, k = j / items, item = j - k * items;
      c0[c] = (uint32_t)(e0 * 3 + item); c1[c] = (uint32_t)(S_TORQUE + k); c2[c] = (uint32_t)step; c3[c] = (uint32_t)((uint64_t)step >> 32);
    }
    uint32_t k0 = key0, k1 = key1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0[c]), lo0 = 0xD2511F53u * c0[c];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2[c]), lo1 = 0xCD9E8D57u * c2[c];
        c0[c] = hi1 ^ c1[c] ^ k0;
        c1[c] = lo1;
        c2[c] = hi0 ^ c3[c] ^ k1;
        c3[c] = lo0;
      }
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int j = j0 + c * nthreads;
      if (j < total) s_u4[j] = make_float4(u01(c0[c]), u01(c1[c]), u01(c2[c]), u01(c3[c]));
    }
  }
}

// The CTA owns TB = env_block consecutive envs.  At the sizes this task runs at the kernel is bound by the length of
// one thread's dependent instruction chain (a warp runs the program exactly once; there are only a few warps per
// scheduler), not by bandwidth, so every env is worked on by POST_ROLES threads ("roles"), each a warp-uniform slice of
// the step:
//   R_BASE     counters, command schedule, derived base state (+ IMU-lag pushes), push / external-force windows,
//              termination, collision                                                    lr:464-481, 509-517; t1:179-247
//   R_BASE_A   base_acc, low_speed, orientation                                          t1:717, 816, 670
//   R_BASE_B   track_vel_hard, tracking_ang_vel, tracking_lin_vel, vel_mismatch_exp      t1:738-790, 726
//   R_JOINT_A  the per-DOF reductions: default_joint_pos, dof_acc, dof_vel, joint_pos, stand_sysmetry
//   R_JOINT_B  feet_distance, knee_distance, and — last, behind the substeps — action_smoothness, torques
//   R_FOOT0/1  one foot each: Euler angles, contact bookkeeping (air time, clearance) and its share of every feet
//              term; foot 1 hands its partial results to foot 0, which adds them in the reference's order
// A role recomputes what it needs of another role's intermediate results (a quaternion rotation is cheaper than a
// barrier).  The unscaled terms meet in shared memory; R_BASE then forms the reward sum in the reference's
// (alphabetical) order while the other roles update the per-term episode sums, and all threads share the reset
// bookkeeping.
// FUSED: the kernel also carries the CTA's envs through the DEC substeps that precede the post-physics phase
// (ti5_fused_step): 3 x WS x TB further threads, WS per (env, four DOFs), run them while the roles
// work; they meet R_JOINT_B on a named barrier in front of the two terms over this step's actions and torques.
constexpr int POST_ROLES = 7;
enum PostRole { R_BASE = 0, R_BASE_A, R_BASE_B, R_JOINT_A, R_JOINT_B, R_FOOT0, R_FOOT1 };

// t1:223-226 the base velocities a push sets: drawn by R_BASE (which applies them) and again, identically, by R_BASE_A
// (whose base_acc term sees the pushed velocities)
struct PushDraw { float fx, fy, tq[3]; };
static __device__ __noinline__ PushDraw push_draw(const Ti5Params& p, const Ti5Rng& r, int64_t step, int e) {
  float u[5];
#pragma unroll
  for (int c = 0; c < 5; ++c)
    u[c] = p.rng_mode == TI5_RNG_PHILOX ? philox_u(p.seed, (uint64_t)step, S_PUSH, e * 5 + c) : r.push[(size_t)e * 5 + c];
  PushDraw d;
  d.fx = affine(p.push_vel_w, p.push_vel_lo, u[0]);
  d.fy = affine(p.push_vel_w, p.push_vel_lo, u[1]);
#pragma unroll
  for (int c = 0; c < 3; ++c) d.tq[c] = affine(p.push_ang_w, p.push_ang_lo, u[2 + c]);
  return d;
}

// what one foot contributes to the feet terms (R_FOOT1 -> R_FOOT0 through shared memory)
enum FootPart { FP_Z = 0, FP_AIR, FP_CLEAR, FP_FORCE, FP_NUMBER, FP_PITCH, FP_STUMBLE, FP_SLIP, FP_STILL, FP_COUNT };
static_assert(FP_COUNT == FOOT_PARTS, "post_tile_bytes reserves FOOT_PARTS floats per env");

// RARE: the options t1_cfg leaves off — commands.heading_command (t1:141-176, 185-188) and commands.sw_switch = False
// (t1:89-90) — behind a compile-time switch: as never-taken run-time branches in the common prologue of the seven roles
// heading_command alone cost the t1 configuration 0.5 us per step (38.3 -> 38.8 us at 8192 envs).
template <bool FUSED, int MAXTB, int WS = 2, bool RARE = false>
__global__ void __launch_bounds__((POST_ROLES + (FUSED ? 3 * WS : 0)) * MAXTB, MAXTB == 32 ? (FUSED && WS == 1 ? 3 : 2) : 1)
post_physics_kernel(const __grid_constant__ Ti5Params p, const __grid_constant__ Ti5Buffers b,
                    const __grid_constant__ Ti5Rng r, const __grid_constant__ PostSrc src,
                    const float* __restrict__ actions_in, int options) {
  chain_trigger();                                        // ti5_reset_observe may become resident
  const int push_last = options & TI5_POST_PUSH_LAST;
  const int N = p.num_envs;
  constexpr int TB = MAXTB;                               // == p.env_block (checked at launch): divisions become shifts
  const int tid = threadIdx.x;
  const int role = tid / TB, le = tid - role * TB;
  const int e0 = blockIdx.x * TB, e = e0 + le, n_tile = min(TB, N - e0);
  const bool is_role = role < POST_ROLES;
  const bool live = e < N && is_role;
  const int role_threads = POST_ROLES * TB;
  Ti5Globals* g = b.globals;
  const uint32_t mask = p.term_mask;
  extern __shared__ __align__(128) unsigned char post_smem[];
  PostTile T;
  T.base = post_smem;
  T.tb = TB;
  T.src = &src;
  T.sums = reinterpret_cast<float*>(post_smem + (size_t)src.off[POST_CHUNKS] * TB);
  T.vals = T.sums + TI5_NUM_TERMS * TB;
  T.bar = reinterpret_cast<uint64_t*>(T.vals + TI5_NUM_TERMS * TB);
  float* s_foot = reinterpret_cast<float*>(T.bar + 2);    // [FP_COUNT][TB] partial results of foot 1
  // the role threads synchronise among themselves; with substep workers in the CTA (FUSED) those run on undisturbed
  auto roles_sync = [&]() {
    if (FUSED) bar_sync(BAR_ROLES, role_threads);
    else __syncthreads();
  };

  // FUSED: [DEC][3 TB] torque-multiplier uniforms, then [DEC][3 TB] lagged action rows, behind the tile
  float4* s_u4 = reinterpret_cast<float4*>(s_foot + FP_COUNT * TB);
  float4* s_old = s_u4 + p.decimation * 3 * TB;
  constexpr int WORKER_THREADS_PER_ENV = 3 * WS;
  constexpr int WORKER_THREADS = FUSED ? WORKER_THREADS_PER_ENV * TB : 0;
  if (FUSED && !is_role) {
    // ======== substep workers: WS threads per (env, four DOFs), coalesced over the CTA's 3 x TB groups; the
    // split index is warp-uniform: the first 3 x TB workers take the substeps 0, WS, ..., the next 1, ... ========
    const int64_t step = g->step_index + 1;
    const int item0 = tid - role_threads, kh = item0 / (3 * TB), item = item0 - kh * 3 * TB;
    const int wl = item / 3, gq = item - wl * 3;
    const bool active = e0 + wl < N;
    probe(b.debug_ts, 2, 0, role_threads);
    SubstepWorker<WS> w;
    if (active) w.setup(p, b, actions_in, step, e0 * 3 + item, e0 + wl, gq, kh, T.at<float>(C_ACT) + wl * D, s_old, item, 3 * TB);
    bar_sync(BAR_DRAWS, role_threads + WORKER_THREADS);    // the uniforms are staged; every worker's old rows are in
    if (active) w.run(p, b, r, step, e0 * 3 + item, e0 + wl, gq, kh, T.at<float>(C_TORQUES) + wl * D, s_u4, s_old, item, 3 * TB);
    probe(b.debug_ts, 2, 1, role_threads);
    bar_arrive(BAR_TORQUES, WORKER_THREADS_PER_ENV * TB + TB);   // R_JOINT_B may read the two tile rows
  }

  // Early mode (chained launch after the substep kernels, full tile): of everything this kernel reads, only `actions`
  // and `torques` come from the substep kernels of this step — every other array is simulator state (untouched while no
  // simulator runs between the kernels) or was last written by the previous step.  All roles but the tail of R_JOINT_B
  // therefore run BEFORE the grid wait; R_JOINT_B then waits for the substeps and pulls the two late arrays.  Nothing
  // stored in front of the wait is read or written by a substep kernel (root_states is: with push_robots the kernel
  // keeps the plain order).  Small grids only (<= 12288 envs): there the CTAs are resident long before the substeps
  // finish (common carve-out, ti5_host.h); on large grids the plain order measured no worse.
  const bool early = !FUSED && (options & TI5_POST_CHAINED) && n_tile == TB && N <= TI5_SMALL_GRID_ENVS && !(p.flags & TI5_F_PUSH_ROBOTS);
  // FUSED: the two arrays come from this CTA's own substep workers, never from memory
  const bool late_by_tma = !FUSED && !early;
  const uint32_t late_bytes = (uint32_t)TB * (uint32_t)(src.rowb[C_ACT] + src.rowb[C_TORQUES]);
  int64_t step = 0, counter = 0;
  bool first_force = false, push_window = false, force_window = false;
  if (is_role) {
    probe(b.debug_ts, 0, 0);
    // ---- stage the CTA's tile of inputs, FIRST: one TMA bulk copy per array, all in flight together ----------
    if (n_tile == TB) {
      if (tid == 0) { mbar_init(T.bar, 1); mbar_init(T.bar + 1, 1); }
      roles_sync();
      // thread k issues bulk copy k of the table; threads 32..59 one episode-sum column each; thread 0 arms the
      // barrier with the byte total (arrival order between the copies and the arm does not matter)
      // (the (TERMS, N) episode-sum columns start 16-byte aligned only when N is a multiple of 4)
      const bool sums_by_tma = (N & 3) == 0;
      if (tid < POST_CHUNKS) {
        if (tid != C_ACT && tid != C_TORQUES && tid != C_CONTACT && tid != C_RIGID)
          tma_load_1d(T.base + (size_t)src.off[tid] * TB, static_cast<const char*>(src.ptr[tid]) + (size_t)e0 * src.rowb[tid],
                      (uint32_t)(TB * src.rowb[tid]), T.bar);
      } else if (sums_by_tma && tid >= 32 && tid < 32 + TI5_NUM_TERMS && (mask & (1u << (tid - 32)))) {
        const int t = tid - 32;
        tma_load_1d(T.sums + (size_t)t * TB, b.episode_sums + (size_t)t * N + e0, (uint32_t)(TB * 4), T.bar);
      }
      if (tid == 0)
        mbar_expect_tx(T.bar, (uint32_t)TB * (uint32_t)(src.off[POST_CHUNKS] - src.rowb[C_CONTACT] - src.rowb[C_RIGID] +
                                                        (sums_by_tma ? 4 * __popc(mask) : 0)) - late_bytes);
      if (!sums_by_tma) {
#pragma unroll 1
        for (int t = 0; t < TI5_NUM_TERMS; ++t)
          if (mask & (1u << t))
            coop_load_n(T.sums + (size_t)t * TB, b.episode_sums + (size_t)t * N + e0, (uint32_t)(TB * 4), tid, role_threads);
      }
      if (late_by_tma) {
        chain_wait();                                     // the substep kernels are done
        if (tid == TB) {                                  // one thread: arm the second barrier and start the two copies
          mbar_expect_tx(T.bar + 1, late_bytes);
          tma_load_1d(T.base + (size_t)src.off[C_ACT] * TB, static_cast<const char*>(src.ptr[C_ACT]) + (size_t)e0 * src.rowb[C_ACT],
                      (uint32_t)(TB * src.rowb[C_ACT]), T.bar + 1);
          tma_load_1d(T.base + (size_t)src.off[C_TORQUES] * TB,
                      static_cast<const char*>(src.ptr[C_TORQUES]) + (size_t)e0 * src.rowb[C_TORQUES],
                      (uint32_t)(TB * src.rowb[C_TORQUES]), T.bar + 1);
        }
      }
    } else {      // partial last tile: byte counts need not be multiples of 16, copy word by word
      if (!FUSED) chain_wait();
#pragma unroll 1
      for (int k = 0; k < POST_CHUNKS; ++k)
        if (k != C_CONTACT && k != C_RIGID && (!FUSED || (k != C_ACT && k != C_TORQUES)))
          coop_load_n(T.base + (size_t)src.off[k] * TB, static_cast<const char*>(src.ptr[k]) + (size_t)e0 * src.rowb[k],
                      (uint32_t)(n_tile * src.rowb[k]), tid, role_threads);
#pragma unroll 1
      for (int t = 0; t < TI5_NUM_TERMS; ++t)
        if (mask & (1u << t))
          coop_load_n(T.sums + (size_t)t * TB, b.episode_sums + (size_t)t * N + e0, (uint32_t)(n_tile * 4), tid, role_threads);
    }
    stage_body_rows(p, b, T.at<float>(C_CONTACT), T.at<float>(C_RIGID), e0, n_tile, tid, role_threads);
    // ---- while the tile is in flight: the step counters and the window predicates (uniform over the grid, t1:193-215)
    step = g->step_index + 1;                             // index of the step in progress
    counter = step + g->common_step_offset;               // common_step_counter after lr:471
    first_force = g->is_first_add_force[step & 1] != 0;
    if (p.flags & TI5_F_PUSH_ROBOTS) {
      int64_t i = fast_div(counter, p.push_update_step);
      if (i >= p.n_push_dur) i = p.n_push_dur - 1;
      push_window = (double)fast_mod(counter, p.push_interval) <= p.push_duration[i];
    }
    if (p.flags & TI5_F_ADD_EXT_FORCE) {
      int64_t i = fast_div(counter, p.add_update_step);
      if (i >= p.n_add_dur) i = p.n_add_dur - 1;
      force_window = (double)fast_mod(counter, p.ext_force_interval) <= p.add_duration[i];
    }
    if (FUSED) {
      // the substep workers' torque-multiplier uniforms, while the tile is in flight
      if (p.flags & TI5_F_RAND_TORQUE) stage_torque_uniforms(p, r, step, e0, 3 * TB, s_u4, tid, role_threads);
      bar_arrive(BAR_DRAWS, role_threads + WORKER_THREADS);
    }
    // a preceding ti5_sample_heights (chained launch) must have completed before this grid does: ti5_reset_observe,
    // which reads the heights, only waits for THIS grid
    if (FUSED && (options & TI5_FUSED_CHAINED)) chain_wait();
    cp_async_wait_all();                                  // this thread's share of the body rows has landed
    roles_sync();
    if (n_tile == TB) {
      mbar_wait(T.bar, 0);
      if (late_by_tma) mbar_wait(T.bar + 1, 0);
    }
  }
  const bool philox = p.rng_mode == TI5_RNG_PHILOX;
  const int dm = p.div_mode;
  // typed views of the tile
  const float* t_root = T.at<float>(C_ROOT);
  const float* t_dof = T.at<float>(C_DOF);
  const float* t_contact = T.at<float>(C_CONTACT);
  const float* t_rigid = T.at<float>(C_RIGID);
  const float* t_act = T.at<float>(C_ACT);
  const float* t_last_act = T.at<float>(C_LAST_ACT);
  const float* t_last_last_act = T.at<float>(C_LAST_LAST_ACT);
  const float* t_last_dof_vel = T.at<float>(C_LAST_DOF_VEL);
  const float* t_torques = T.at<float>(C_TORQUES);
  const float* t_ref = T.at<float>(C_REF);
  const float* t_last_root_vel = T.at<float>(C_LAST_ROOT_VEL);
  const float* t_cmd = T.at<float>(C_CMD);
  const float* t_gait_start = T.at<float>(C_GAIT_START);
  const float* t_air = T.at<float>(C_AIR);
  const float* t_feet_h = T.at<float>(C_FEET_H);
  const float* t_last_z = T.at<float>(C_LAST_Z);
  const int64_t* t_ep_len = T.at<int64_t>(C_EP_LEN);
  const int64_t* t_phase_len = T.at<int64_t>(C_PHASE_LEN);
  const int32_t* t_gait_time = T.at<int32_t>(C_GAIT_TIME);
  const uint8_t* t_last_contacts = T.at<uint8_t>(C_LAST_CONTACTS);
  probe(b.debug_ts, 0, 1);
  bool reset = false, time_out = false;
  auto put = [&](int t, float v) { T.vals[t * TB + le] = v; };
  // what the two foot roles carry across their barrier
  float part[FP_COUNT], stance[2] = {0.0f, 0.0f};
  bool stand_foot = false;

  if (live) {
    // ======== common prologue (every role, same values): command schedule, stand flag =====================
    // t1:183-184 phase counter and gait-schedule command resampling (pass 0)
    const int64_t ep_len = t_ep_len[le] + 1;                                  // lr:469
    int64_t phase_len = t_phase_len[le] + 1;
    float4 cmd = reinterpret_cast<const float4*>(t_cmd)[le];
    for (int gi = 0; gi < p.num_gaits; ++gi) {
      if (ep_len != (int64_t)t_gait_time[le * p.num_gaits + gi]) continue;
      const int kind = p.gait_kind[gi];
      float u[3];
      if (philox) {                                       // one Philox call for the three components
        const float4 a = philox_u4(p.seed, (uint64_t)step, S_CMD + gi, e);
        u[0] = a.x; u[1] = a.y; u[2] = a.z;
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) u[c] = r.cmd[((size_t)(0 * p.num_gaits + gi) * N + e) * 3 + c];
      }
      const bool mx = kind == TI5_GAIT_WALK_SAGITTAL || kind == TI5_GAIT_WALK_OMNI;
      const bool my = kind == TI5_GAIT_WALK_LATERAL || kind == TI5_GAIT_WALK_OMNI;
      const bool mz = kind == TI5_GAIT_ROTATE || kind == TI5_GAIT_WALK_OMNI;
      const double (*cr)[2] = g->cmd_range[step & 1];    // ranges before this step's curriculum update (lr:537 runs later)
      cmd.x = mx ? affine((float)(cr[0][1] - cr[0][0]), (float)cr[0][0], u[0]) : 0.0f;
      cmd.y = my ? affine((float)(cr[1][1] - cr[1][0]), (float)cr[1][0], u[1]) : 0.0f;
      // heading mode (t1:141-176): the third draw is the heading target; the yaw rate follows below
      if (RARE && (p.flags & TI5_F_HEADING_COMMAND)) cmd.w = mz ? affine(p.heading_w, p.heading_lo, u[2]) : 0.0f;
      else cmd.z = mz ? affine((float)(cr[2][1] - cr[2][0]), (float)cr[2][0], u[2]) : 0.0f;
    }
    const float bq[4] = {t_root[le * RB + 3], t_root[le * RB + 4], t_root[le * RB + 5], t_root[le * RB + 6]};
    if (RARE && (p.flags & TI5_F_HEADING_COMMAND)) cmd.z = heading_yaw_rate(bq, cmd.w);       // t1:185-188, every env, every step
    const float cmd_norm = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y + cmd.z * cmd.z);
    const bool stand = cmd_norm <= p.stand_threshold;
    const bool no_sw = RARE && (p.flags & TI5_F_NO_SW_SWITCH);                 // t1:89-90: phase from the episode counter
    if (stand && !no_sw) phase_len = 0;                   // t1:86 side effect: standing envs restart the phase
    const float* qrow = t_dof + (size_t)le * 2 * D;                           // interleaved (q, qd)

    if (role == R_BASE) {
      // ================================ the base: state ================================================
      float root[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) root[i] = t_root[le * RB + i];
      const float* tf = t_contact + le * CON_STRIDE + 0;
      const float term_force = sqrtf(tf[0] * tf[0] + tf[1] * tf[1] + tf[2] * tf[2]);
      const float* pf = t_contact + le * CON_STRIDE + 3;
      const float pen_force = sqrtf(pf[0] * pf[0] + pf[1] * pf[1] + pf[2] * pf[2]);
      b.episode_length_buf[e] = ep_len;
      b.phase_length_buf[e] = phase_len;
      reinterpret_cast<float4*>(b.commands)[e] = cmd;
      // lr:475-479 derived base state
      const V3 lin = quat_rotate_inverse(bq, V3{root[7], root[8], root[9]});
      const V3 ang = quat_rotate_inverse(bq, V3{root[10], root[11], root[12]});
      const V3 grav = quat_rotate_inverse(bq, V3{0.0f, 0.0f, -1.0f});
      float eul[3];
      euler_xyz(bq, eul);
      reinterpret_cast<float4*>(b.base_quat)[e] = make_float4(bq[0], bq[1], bq[2], bq[3]);
      b.base_lin_vel[e * 3 + 0] = lin.x; b.base_lin_vel[e * 3 + 1] = lin.y; b.base_lin_vel[e * 3 + 2] = lin.z;
      b.base_ang_vel[e * 3 + 0] = ang.x; b.base_ang_vel[e * 3 + 1] = ang.y; b.base_ang_vel[e * 3 + 2] = ang.z;
      b.projected_gravity[e * 3 + 0] = grav.x; b.projected_gravity[e * 3 + 1] = grav.y; b.projected_gravity[e * 3 + 2] = grav.z;
#pragma unroll
      for (int i = 0; i < 3; ++i) b.base_euler_xyz[e * 3 + i] = eul[i];
      if (FUSED && (p.flags & TI5_F_ADD_IMU_LAG)) {
        // lr:428-434 for every substep: without a simulator in between the base does not move, so each push is the
        // derived state of this step (same quaternion, same arithmetic as the substep kernel's pieces)
        int slot = (int)fast_mod((step - 1) * p.decimation, p.imu_lag_len);
#pragma unroll 1
        for (int k = 0; k < p.decimation; ++k) {
          float* row = b.imu_ring + ((size_t)slot * N + e) * 6;
          row[0] = ang.x; row[1] = ang.y; row[2] = ang.z; row[3] = eul[0]; row[4] = eul[1]; row[5] = eul[2];
          slot = slot + 1 == p.imu_lag_len ? 0 : slot + 1;
        }
      } else if (push_last && (p.flags & TI5_F_ADD_IMU_LAG)) {               // fused IMU-lag push of the last substep
        const int64_t j = (step - 1) * p.decimation + (p.decimation - 1);
        float* row = b.imu_ring + ((size_t)ring_slot(j, p.imu_lag_len) * N + e) * 6;
        row[0] = ang.x; row[1] = ang.y; row[2] = ang.z; row[3] = eul[0]; row[4] = eul[1]; row[5] = eul[2];
      }
      // t1:193-203, 217-231 push window: overwrite the base velocity
      if (p.flags & TI5_F_PUSH_ROBOTS) {
        float tq[3] = {0.0f, 0.0f, 0.0f};
        if (push_window) {
          const PushDraw d = push_draw(p, r, step, e);
          float* rw = b.root_states + (size_t)e * RB;
          rw[7] = d.fx; rw[8] = d.fy; rw[10] = d.tq[0]; rw[11] = d.tq[1]; rw[12] = d.tq[2];
          b.rand_push_force[e * 3 + 0] = d.fx;
          b.rand_push_force[e * 3 + 1] = d.fy;
          tq[0] = d.tq[0]; tq[1] = d.tq[1]; tq[2] = d.tq[2];
        } else {
          b.rand_push_force[e * 3 + 0] = 0.0f; b.rand_push_force[e * 3 + 1] = 0.0f; b.rand_push_force[e * 3 + 2] = 0.0f;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) b.rand_push_torque[e * 3 + c] = tq[c];
      }
      // t1:205-215, 233-247 external force window
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
            const float sc = stand ? 1.0f : 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              af[c] = b.ext_forces[e * 3 + c] * sc;
              at[c] = b.ext_torques[e * 3 + c] * sc;
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) { b.ext_forces[e * 3 + c] = 0.0f; b.ext_torques[e * 3 + c] = 0.0f; }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          b.applied_force[(size_t)e * p.applied_stride + c] = af[c];
          b.applied_torque[(size_t)e * p.applied_stride + c] = at[c];
        }
      }
      // lr:509-517 termination
      time_out = ep_len > p.max_episode_length;
      reset = (term_force > 1.0f) || time_out;
      b.time_out_buf[e] = time_out ? 1 : 0;
      b.reset_buf[e] = reset ? 1 : 0;
      // the arrival-order work list for the history clear (ti5_reset_observe), here rather than in the epilogue
      if (reset) b.reset_list[atomicAdd(&g->n_listed[step & 1], 1)] = e;
      if (mask & (1u << T_COLLISION)) put(T_COLLISION, 1.0f * (pen_force > 0.1f ? 1.0f : 0.0f));   // t1:870-875
    } else if (role == R_BASE_A) {
      // ================================ the base: acceleration, speed, orientation ====================
      if (mask & (1u << T_BASE_ACC)) {                      // t1:717-724, on the velocities a push may just have set
        float v[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = t_root[le * RB + 7 + i];
        if ((p.flags & TI5_F_PUSH_ROBOTS) && push_window) {
          const PushDraw d = push_draw(p, r, step, e);
          v[0] = d.fx; v[1] = d.fy; v[3] = d.tq[0]; v[4] = d.tq[1]; v[5] = d.tq[2];
        }
        float sq = 0.0f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const float d = t_last_root_vel[le * 6 + i] - v[i];
          sq += d * d;
        }
        put(T_BASE_ACC, expf_call(-sqrtf(sq) * 3.0f));
      }
      if (mask & (1u << T_LOW_SPEED)) {                     // t1:816-847 (appendix A13)
        const V3 lin = quat_rotate_inverse(bq, V3{t_root[le * RB + 7], t_root[le * RB + 8], t_root[le * RB + 9]});
        const float av = fabsf(lin.x), ac = fabsf(cmd.x);
        const bool slow = av < 0.5f * ac, fast = av > 1.2f * ac;
        float v = 0.0f;
        if (slow) v = -1.0f;
        if (fast) v = 0.0f;
        if (!(slow || fast)) v = 1.2f;
        if (signf(lin.x) != signf(cmd.x)) v = -2.0f;
        put(T_LOW_SPEED, v * (ac > 0.05f ? 1.0f : 0.0f));
      }
      if (mask & (1u << T_ORIENTATION)) {                   // t1:670-677
        const V3 grav = quat_rotate_inverse(bq, V3{0.0f, 0.0f, -1.0f});
        const float a = expf_call(-(fabsf(euler_roll(bq)) + fabsf(euler_pitch(bq))) * 10.0f);
        const float bb = expf_call(-sqrtf(grav.x * grav.x + grav.y * grav.y) * 20.0f);
        put(T_ORIENTATION, (a + bb) / 2.0f);
      }
    } else if (role == R_BASE_B) {
      // ================================ the base: velocity tracking ====================================
      const V3 lin = quat_rotate_inverse(bq, V3{t_root[le * RB + 7], t_root[le * RB + 8], t_root[le * RB + 9]});
      const V3 ang = quat_rotate_inverse(bq, V3{t_root[le * RB + 10], t_root[le * RB + 11], t_root[le * RB + 12]});
      const float ex = cmd.x - lin.x, ey = cmd.y - lin.y, ew = cmd.z - ang.z;
      if (mask & (1u << T_TRACK_VEL_HARD)) {                // t1:738-758
        const float le_ = sqrtf(ex * ex + ey * ey);
        const float ae = fabsf(ew);
        put(T_TRACK_VEL_HARD, (expf_call(-le_ * 10.0f) + expf_call(-ae * 10.0f)) / 2.0f - 0.2f * (le_ + ae));
      }
      if (mask & (1u << T_TRACKING_ANG_VEL))                // t1:776-790
        put(T_TRACKING_ANG_VEL, expf_call(stand ? -fabsf(ew) * p.tracking_sigma * 2.0f : -(ew * ew) * p.tracking_sigma));
      if (mask & (1u << T_TRACKING_LIN_VEL))                // t1:760-774
        put(T_TRACKING_LIN_VEL, expf_call(stand ? -(fabsf(ex) + fabsf(ey)) * p.tracking_sigma * 2.0f
                                                : -(ex * ex + ey * ey) * p.tracking_sigma));
      if (mask & (1u << T_VEL_MISMATCH_EXP)) {              // t1:726-736
        const float a = expf_call(-(lin.z * lin.z) * 10.0f);
        const float bb = expf_call(-sqrtf(ang.x * ang.x + ang.y * ang.y) * 5.0f);
        put(T_VEL_MISMATCH_EXP, (a + bb) / 2.0f);
      }
    } else if (role == R_JOINT_A) {
      // ================================ the joints: per-DOF reductions =================================
      if (!FUSED && push_last && (p.flags & TI5_F_ADD_DOF_LAG)) {            // fused DOF-lag push of the last substep
        const int64_t j = (step - 1) * p.decimation + (p.decimation - 1);
        float* row = b.dof_ring + ((size_t)ring_slot(j, p.dof_lag_len) * N + e) * (2 * D);
#pragma unroll 1
        for (int i = 0; i < D; ++i) { row[i] = qrow[2 * i]; row[D + i] = qrow[2 * i + 1]; }
      }
      const float* ldv = t_last_dof_vel + le * D;
      const float* ref = t_ref + le * D;
      // one pass over the 12 DOFs feeds every per-DOF reduction that needs joint state only (each sum keeps its own
      // DOF order)
      float s_dq = 0.0f, s_acc = 0.0f, s_vel = 0.0f, s_jp = 0.0f;
#pragma unroll 1
      for (int i = 0; i < D; ++i) {
        const float qi = qrow[2 * i], qdi = qrow[2 * i + 1];
        const float dq = qi - p.default_dof_pos[i];
        s_dq += dq * dq;
        const float ac = sdiv(ldv[i] - qdi, p.dt, dm);
        s_acc += ac * ac;
        s_vel += qdi * qdi;
        const float dj = qi - (stand ? p.default_dof_pos[i] : ref[i]);       // ref_dof_pos of the PREVIOUS step (A3)
        s_jp += dj * dj;
      }
      auto dq0 = [&](int i) { return qrow[2 * i] - p.default_dof_pos[i]; };
      if (mask & (1u << T_DEFAULT_JOINT_POS)) {             // t1:686-703
        const float l = sqrtf((dq0(0) * dq0(0) + dq0(1) * dq0(1)) + dq0(5) * dq0(5));
        const float rr = sqrtf((dq0(6) * dq0(6) + dq0(7) * dq0(7)) + dq0(11) * dq0(11));
        const float yr = clampf((l + rr) - 0.1f, 0.0f, 50.0f);
        put(T_DEFAULT_JOINT_POS, expf_call(-yr * 100.0f) - 0.01f * sqrtf(s_dq));
      }
      if (mask & (1u << T_DOF_ACC)) put(T_DOF_ACC, s_acc);       // t1:863-868
      if (mask & (1u << T_DOF_VEL)) put(T_DOF_VEL, s_vel);       // t1:856-861
      if (mask & (1u << T_JOINT_POS)) {                     // t1:576-596
        const float n = sqrtf(s_jp);
        const float v = expf_call(-2.0f * n) - 0.2f * clampf(n, 0.0f, 0.5f);
        put(T_JOINT_POS, stand ? 1.0f : v);
      }
      if (mask & (1u << T_STAND_SYSMETRY)) {                // t1:917-925
        float sq = 0.0f;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const float d = qrow[2 * i] - qrow[2 * (5 + i)];
          sq += d * d;
        }
        put(T_STAND_SYSMETRY, stand ? expf_call(-sq) : 0.0f);
      }
    } else if (role == R_JOINT_B) {
      // ================================ the joints: distances (actions and torques follow below) ======
      const float* f0 = t_rigid + le * RIG_STRIDE + 0 * RB;
      const float* f1 = t_rigid + le * RIG_STRIDE + 1 * RB;
      const float* k0 = t_rigid + le * RIG_STRIDE + 2 * RB;
      const float* k1 = t_rigid + le * RIG_STRIDE + 3 * RB;
      if (mask & (1u << T_FEET_DISTANCE))                   // t1:599-612
        put(T_FEET_DISTANCE, pair_distance_reward(f0[0], f0[1], f1[0], f1[1], p.foot_min_dist, p.foot_max_dist));
      if (mask & (1u << T_KNEE_DISTANCE))                   // t1:615-628
        put(T_KNEE_DISTANCE, pair_distance_reward(k0[0], k0[1], k1[0], k1[1], p.knee_min_dist, p.knee_max_dist));
    } else {
      // ================================ one foot (f = 0: R_FOOT0, f = 1: R_FOOT1) =====================
      const int f = role - R_FOOT0;
      // gait phase and stance mask (t1:80-107)
      const float phase = (py_mod1(sdiv((float)(no_sw ? ep_len : phase_len) * p.dt, p.cycle_time, dm)) + t_gait_start[le]) *
                          (stand && !no_sw ? 0.0f : 1.0f);
      const float sin_pos = sinf(TWO_PI_F * phase);
      stance[0] = sin_pos >= 0.0f ? 1.0f : 0.0f;
      stance[1] = sin_pos < 0.0f ? 1.0f : 0.0f;
      if (fabsf(sin_pos) < 0.1f) stance[0] = stance[1] = 1.0f;
      stand_foot = stand;
      const float* cf = t_contact + le * CON_STRIDE + (2 + f) * 3;
      const bool contact = cf[2] > 5.0f;
      const float* rs = t_rigid + le * RIG_STRIDE + f * RB;
      const float fq[4] = {rs[3], rs[4], rs[5], rs[6]};
      float fe[3];
      euler_xyz(fq, fe);                                                       // lr:480-481
#pragma unroll
      for (int i = 0; i < 3; ++i) b.feet_euler_xyz[e * 6 + 3 * f + i] = fe[i];
#pragma unroll
      for (int i = 0; i < FP_COUNT; ++i) part[i] = 0.0f;
      part[FP_Z] = rs[2];
      part[FP_PITCH] = fe[1];
      if (mask & (1u << T_FEET_AIR_TIME)) {                 // t1:642-657 (appendix A6, A8)
        const bool tiny = cmd_norm < 0.05f;
        const float st = tiny ? 1.0f : stance[f];
        const bool filt = contact || (st != 0.0f) || (t_last_contacts[le * 2 + f] != 0);
        b.contact_filt[e * 2 + f] = filt ? 1 : 0;
        b.last_contacts[e * 2 + f] = contact ? 1 : 0;
        float air = t_air[le * 2 + f];
        const float first = (air > 0.0f && filt) ? 1.0f : 0.0f;
        air += p.dt;
        part[FP_AIR] = clampf(air, 0.0f, 0.5f) * first;
        b.feet_air_time[e * 2 + f] = air * (filt ? 0.0f : 1.0f);
      }
      if (mask & (1u << T_FEET_CLEARANCE)) {                // t1:793-814 (appendix A9)
        const float z = rs[2];
        const float h = t_feet_h[le * 2 + f] + (z - t_last_z[le * 2 + f]);
        b.last_feet_z[e * 2 + f] = z;
        const float swing = 1.0f - stance[f];
        const float hit = (h > p.target_feet_height && h < p.target_feet_height_max) ? 1.0f : 0.0f;
        part[FP_CLEAR] = hit * swing;
        b.feet_height[e * 2 + f] = h * (contact ? 0.0f : 1.0f);
      }
      if (mask & (1u << T_FEET_CONTACT_FORCES)) {           // t1:679-684
        const float n = sqrtf((cf[0] * cf[0] + cf[1] * cf[1]) + cf[2] * cf[2]);
        part[FP_FORCE] = clampf(n - p.max_contact_force, 0.0f, 400.0f);
      }
      if (mask & (1u << T_FEET_CONTACT_NUMBER)) {           // t1:659-668 (appendix A14)
        const float st = stand ? 1.0f : stance[f];
        part[FP_NUMBER] = ((contact ? 1.0f : 0.0f) == st) ? 1.0f : -0.3f;
      }
      if (mask & (1u << T_FEET_STUMBLE))                    // t1:937-940
        part[FP_STUMBLE] = sqrtf(cf[0] * cf[0] + cf[1] * cf[1]) > 5.0f * fabsf(cf[2]) ? 1.0f : 0.0f;
      if (mask & (1u << T_FOOT_SLIP))                       // t1:630-640 (appendix A10): rigid_state[..., 10:12]
        part[FP_SLIP] = sqrtf(sqrtf(rs[10] * rs[10] + rs[11] * rs[11])) * (contact ? 1.0f : 0.0f);
      if (f == 1) {
        if (mask & (1u << T_STAND_STILL)) {                 // t1:899-915 (appendix A12), the joint part:
          // dof_idx [0,1,2,3,5,6,7,8] with weights [2,2,1,1,1,2,2,1]; foot 0 adds the two foot pitches with weight 1
          float sq = 0.0f;
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            const int j = i < 4 ? i : i + 1;
            const float w = (i < 2 || i == 5 || i == 6) ? 2.0f : 1.0f;
            const float er = (qrow[2 * j] - p.default_dof_pos[j]) * w;
            sq += er * er;
          }
          part[FP_STILL] = sq;
        }
#pragma unroll
        for (int i = 0; i < FP_COUNT; ++i) s_foot[i * TB + le] = part[i];
      }
    }
  }
  // ---- foot 0 adds up both feet, in the reference's order (0 + foot 0 + foot 1).  The barrier instructions sit outside
  // every `live` branch: a warp executes them as a whole (partial tiles have idle lanes) -------------------------
  if (is_role && role == R_FOOT1) bar_arrive(BAR_FEET, 2 * TB);
  if (is_role && role == R_FOOT0) {
    bar_sync(BAR_FEET, 2 * TB);
    if (live) {
      const bool stand = stand_foot;
      float o[FP_COUNT];
#pragma unroll
      for (int i = 0; i < FP_COUNT; ++i) o[i] = s_foot[i * TB + le];
      if (mask & (1u << T_BASE_HEIGHT)) {                 // t1:706-715
        const float ground = (part[FP_Z] * stance[0] + o[FP_Z] * stance[1]) / (stance[0] + stance[1]);
        const float h = t_root[le * RB + 2] - (ground - 0.05f);
        put(T_BASE_HEIGHT, expf_call(-fabsf(h - p.base_height_target) * 100.0f));
      }
      if (mask & (1u << T_FEET_AIR_TIME)) put(T_FEET_AIR_TIME, (0.0f + part[FP_AIR]) + o[FP_AIR]);
      if (mask & (1u << T_FEET_CLEARANCE)) put(T_FEET_CLEARANCE, (0.0f + part[FP_CLEAR]) + o[FP_CLEAR]);
      if (mask & (1u << T_FEET_CONTACT_FORCES)) put(T_FEET_CONTACT_FORCES, (0.0f + part[FP_FORCE]) + o[FP_FORCE]);
      if (mask & (1u << T_FEET_CONTACT_NUMBER)) put(T_FEET_CONTACT_NUMBER, ((0.0f + part[FP_NUMBER]) + o[FP_NUMBER]) / 2.0f);
      if (mask & (1u << T_FEET_ROTATION)) {               // t1:926-935 (appendix A11)
        const float rot = part[FP_PITCH] * part[FP_PITCH] + o[FP_PITCH] * o[FP_PITCH];
        const float x = rot / 1.0f;
        put(T_FEET_ROTATION, 1.0f * expf_call(-(x * x)));
      }
      if (mask & (1u << T_FEET_STUMBLE)) put(T_FEET_STUMBLE, (part[FP_STUMBLE] != 0.0f || o[FP_STUMBLE] != 0.0f) ? 1.0f : 0.0f);
      if (mask & (1u << T_FOOT_SLIP)) put(T_FOOT_SLIP, (0.0f + part[FP_SLIP]) + o[FP_SLIP]);
      if (mask & (1u << T_STAND_STILL)) {
        float sq = o[FP_STILL];
        const float e0p = part[FP_PITCH] * 1.0f, e1p = o[FP_PITCH] * 1.0f;
        sq += e0p * e0p;
        sq += e1p * e1p;
        put(T_STAND_STILL, stand ? expf_call(-sq) : 0.0f);
      }
    }
  }
  // ---- R_JOINT_B, last: the two terms over this step's actions and torques ---------------------------------------
  if (is_role && role == R_JOINT_B) {
    if (FUSED) bar_sync(BAR_TORQUES, WORKER_THREADS_PER_ENV * TB + TB);     // the CTA's substep workers have left both rows in the tile
    if (live) {
      const float* act = t_act + le * D;
      const float* la = t_last_act + le * D;
      const float* lla = t_last_last_act + le * D;
      const float* tau = t_torques + le * D;
      if (early) {
        chain_wait();                                       // the substep kernels are done
        // the two late rows of this env, straight from the L2 into the env's own tile rows (shorter than a bulk-copy
        // round trip for 96 bytes per thread)
        const float4* ga = reinterpret_cast<const float4*>(b.actions + (size_t)e * D);
        const float4* gt = reinterpret_cast<const float4*>(b.torques + (size_t)e * D);
        const float4 a0 = ga[0], a1 = ga[1], a2 = ga[2], t0 = gt[0], t1 = gt[1], t2 = gt[2];
        float4* sa = reinterpret_cast<float4*>(const_cast<float*>(act));
        float4* st = reinterpret_cast<float4*>(const_cast<float*>(tau));
        sa[0] = a0; sa[1] = a1; sa[2] = a2; st[0] = t0; st[1] = t1; st[2] = t2;
      }
      float s_d1 = 0.0f, s_d2 = 0.0f, s_abs = 0.0f, s_tau = 0.0f;
#pragma unroll 1
      for (int i = 0; i < D; ++i) {
        const float a = act[i], l = la[i];
        const float d1 = (l - a) * 1.0f;
        const float d2 = ((a + lla[i]) - 2.0f * l) * 1.0f;
        s_d1 += d1 * d1;
        s_d2 += d2 * d2;
        s_abs += fabsf(a * 1.0f);
        s_tau += tau[i] * tau[i];
      }
      if (mask & (1u << T_ACTION_SMOOTHNESS)) put(T_ACTION_SMOOTHNESS, (s_d1 + s_d2) + 0.05f * s_abs);   // t1:877-892
      if (mask & (1u << T_TORQUES)) put(T_TORQUES, s_tau);       // t1:849-854
    }
  }
  probe(b.debug_ts, 0, 2);
  probe(b.debug_ts, 0, 6, R_JOINT_B * TB);      // R_JOINT_B done
  probe(b.debug_ts, 0, 7, R_FOOT0 * TB);        // R_FOOT0 done
  __syncthreads();
  probe(b.debug_ts, 0, 3);

  // ---- lr:654-680: reward sum in alphabetical term order, per-term episode sums, clip at zero -----------
  // The per-term episode sums are independent of one another: the other roles take every sixth term each while R_BASE
  // runs the ordered sum (the scaled term is recomputed there: same product, same bits).
  if (live && role > 0) {
#pragma unroll 1
    for (int t = role - 1; t < TI5_NUM_TERMS; t += POST_ROLES - 1) {
      if (!(mask & (1u << t)) || t == T_TERMINATION) continue;
      const float sc = T.vals[t * TB + le] * p.reward_scale[t];
      const float acc = T.sums[t * TB + le] + sc;
      T.sums[t * TB + le] = acc;
      b.episode_sums[(size_t)t * N + e] = acc;
      if (b.reward_terms) b.reward_terms[(size_t)t * N + e] = sc;
    }
  }
  if (live && role == 0) {
    // unrolled: the loads and products of all terms are independent, only the additions form a chain (in the
    // reference's order); terms without a scale add nothing
    float sc[TI5_NUM_TERMS];
#pragma unroll
    for (int t = 0; t < TI5_NUM_TERMS; ++t) sc[t] = T.vals[t * TB + le] * p.reward_scale[t];
    float rew = 0.0f;
#pragma unroll
    for (int t = 0; t < TI5_NUM_TERMS; ++t)
      if (t != T_TERMINATION && (mask & (1u << t))) rew += sc[t];
    if ((p.flags & TI5_F_ONLY_POSITIVE) && rew < 0.0f) rew = 0.0f;     // clip(min=0); NaN passes
    if (mask & (1u << T_TERMINATION)) {                   // lr:677-680, t1:894-896: added after the clip
      const float sc = ((reset && !time_out) ? 1.0f : 0.0f) * p.reward_scale[T_TERMINATION];
      rew += sc;
      const float acc = T.sums[T_TERMINATION * TB + le] + sc;
      T.sums[T_TERMINATION * TB + le] = acc;
      b.episode_sums[(size_t)T_TERMINATION * N + e] = acc;
      if (b.reward_terms) b.reward_terms[(size_t)T_TERMINATION * N + e] = sc;
    }
    b.rew_buf[e] = rew;
  }
  probe(b.debug_ts, 0, 4);
  // t1:205-215: `is_first_add_force` of the next step (double-buffered by step parity: no CTA of this grid reads it)
  if (blockIdx.x == 0 && tid == 0) {
    g->step_now = step;                              // read by ti5_reset_observe (nobody in this grid reads it)
    g->n_listed[(step + 1) & 1] = 0;                 // the next step's work-list counter (this grid fills [step & 1])
    if (p.flags & TI5_F_ADD_EXT_FORCE) g->is_first_add_force[(step + 1) & 1] = force_window ? 0 : 1;
  }
  // (every thread of the CTA calls the bookkeeping; `step` is only used by the threads that hold a flag: role 0)
  reset_bookkeeping(p, b, reset, e, le, T.sums, TB, step, true);
  probe(b.debug_ts, 0, 5);
}

// Bookkeeping for an explicit `reset_idx(env_ids)` call (lr:450-455 `reset()`): the caller has
// written the mask into reset_buf; no physics, no rewards.
__global__ void __launch_bounds__(128)
reset_bookkeeping_kernel(const __grid_constant__ Ti5Params p, const __grid_constant__ Ti5Buffers b) {
  const int N = p.num_envs;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const bool reset = e < N && b.reset_buf[e] != 0;
  // an explicit reset happens between steps: the scatter that follows works at the count of completed steps
  const int64_t step = b.globals->step_index;
  if (e == 0) b.globals->step_now = step;
  reset_bookkeeping(p, b, reset, e, threadIdx.x, nullptr, 0, step, false);
}

}  // namespace ti5

using namespace ti5;

extern "C" int ti5_reset_bookkeeping(const Ti5Params* p, const Ti5Buffers* b, void* stream) {
  TI5_CHECK_ARGS(p && b && p->num_envs > 0);
  TI5_CHECK_ARGS(p->env_block == 32 || p->env_block == 64 || p->env_block == 128);
  const int blocks = (p->num_envs + p->env_block - 1) / p->env_block;
  cudaMemsetAsync(reinterpret_cast<char*>(b->globals) + offsetof(Ti5Globals, n_listed), 0, 2 * sizeof(int32_t), (cudaStream_t)stream);
  reset_bookkeeping_kernel<<<blocks, p->env_block, 0, (cudaStream_t)stream>>>(*p, *b);
  return ti5_check_launch("ti5_reset_bookkeeping");
}

static int launch_post(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, const float* actions_in, int options,
                       bool fused, void* stream, const char* what) {
  TI5_CHECK_ARGS(p && b && p->num_envs > 0 && (options & ~3) == 0);
  TI5_CHECK_ARGS(p->env_block == 32 || p->env_block == 64 || (!fused && p->env_block == 128));
  TI5_CHECK_ARGS(p->num_gaits >= 0 && p->num_gaits <= TI5_MAX_GAITS);
  TI5_CHECK_ARGS(p->rng_mode == TI5_RNG_PHILOX || (r && r->cmd));
  // the per-substep re-draw of the action lag (off in t1_cfg) exists in the unfused substep kernels only
  TI5_CHECK_ARGS(!(fused && (p->flags & TI5_F_LAG_PERSTEP)));
  TI5_CHECK_ARGS((p->term_mask & (1u << T_DOF_VEL_LIMITS)) == 0);   // the reference term reads a cfg field t1 lacks
  TI5_CHECK_ARGS(!(p->flags & TI5_F_ADD_EXT_FORCE) || p->applied_stride >= 3);
  Ti5Rng rr = r ? *r : Ti5Rng{};
  const int blocks = (p->num_envs + p->env_block - 1) / p->env_block;
  const PostSrc src = make_post_src(*p, *b);
  const size_t smem = post_tile_bytes(p->env_block, src.off[POST_CHUNKS], fused ? p->decimation : 0);
  // large grids: one worker thread per (env, four DOFs) (see WS); TI5_WORKER_SPLIT=1|2 overrides
  static const int forced_ws = getenv("TI5_WORKER_SPLIT") ? atoi(getenv("TI5_WORKER_SPLIT")) : 0;
  const bool rare = (p->flags & (TI5_F_HEADING_COMMAND | TI5_F_NO_SW_SWITCH)) != 0;     // RARE builds exist with two workers only
  const int ws = !fused || rare ? 2 : forced_ws ? forced_ws : (ti5_small_grid(p) ? 2 : 1);
  auto kernel = fused ? (p->env_block == 32 ? (ws == 1 ? post_physics_kernel<true, 32, 1> : post_physics_kernel<true, 32, 2>)
                                            : (ws == 1 ? post_physics_kernel<true, 64, 1> : post_physics_kernel<true, 64, 2>))
                      : (p->env_block == 32 ? post_physics_kernel<false, 32>
                                            : p->env_block == 64 ? post_physics_kernel<false, 64> : post_physics_kernel<false, 128>);
  if (rare)
    kernel = fused ? (p->env_block == 32 ? post_physics_kernel<true, 32, 2, true> : post_physics_kernel<true, 64, 2, true>)
                   : (p->env_block == 32 ? post_physics_kernel<false, 32, 2, true>
                                         : p->env_block == 64 ? post_physics_kernel<false, 64, 2, true> : post_physics_kernel<false, 128, 2, true>);
  if (!ti5_ensure_smem(kernel, smem)) {
    ti5_set_error("%s: %zu bytes of shared memory per CTA not available", what, smem);
    return TI5_ECUDA;
  }
  ti5_set_carveout(kernel, ti5_small_grid(p));
  const int threads = (POST_ROLES + (fused ? 3 * ws : 0)) * p->env_block;
  (void)ti5_launch(kernel, dim3(blocks), dim3(threads), smem, stream, (options & TI5_POST_CHAINED) != 0, *p, *b, rr, src,
                   actions_in, options);
  return ti5_check_launch(what);
}

extern "C" int ti5_post_physics(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, int options, void* stream) {
  return launch_post(p, b, r, nullptr, options, false, stream, "ti5_post_physics");
}

extern "C" int ti5_fused_step(const Ti5Params* p, const Ti5Buffers* b, const Ti5Rng* r, const float* actions_in, int options,
                              void* stream) {
  TI5_CHECK_ARGS(actions_in != nullptr && (options & ~TI5_FUSED_CHAINED) == 0);
  TI5_CHECK_ARGS(p && p->decimation >= 1 && p->decimation <= 16);       // Philox sites S_TORQUE + k stay below S_CMD
  TI5_CHECK_ARGS(p->rng_mode == TI5_RNG_PHILOX || !(p->flags & TI5_F_RAND_TORQUE) || (r && r->torque));
  return launch_post(p, b, r, actions_in, options, true, stream, "ti5_fused_step");
}
