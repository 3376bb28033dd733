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
static __device__ __noinline__ void reset_bookkeeping(const Ti5Params& p, const Ti5Buffers& b, bool reset, int e, int le,
                                                      const float* sums, int tb, int64_t step) {
  __shared__ int s_warp[32];
  __shared__ float s_red[12][TI5_NUM_TERMS];
  const int N = p.num_envs, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  if (reset) b.reset_list[atomicAdd(&b.globals->n_listed[step & 1], 1)] = e;
  const BlockRank br = block_rank(reset, s_warp);
  if (tid == 0) b.block_counts[blockIdx.x] = br.total;
  if (br.total == 0) return;
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
  const int rowb[POST_CHUNKS] = {RB * 4, 2 * D * 4, NB * 3 * 4, NB * RB * 4, D * 4, D * 4, D * 4, D * 4, D * 4, D * 4, 6 * 4, 4 * 4,
                                 4, 2 * 4, 2 * 4, 2 * 4, 8, 8, p.num_gaits * 4, 2};
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

__host__ __device__ inline size_t post_tile_bytes(int tb, int per_env_bytes) {
  return (size_t)tb * per_env_bytes + 2 * (size_t)TI5_NUM_TERMS * tb * 4 + 16;
}

// The CTA owns TB = env_block consecutive envs and runs POST_ROLES x TB threads: every env is worked on by
// three threads ("roles"), each evaluating a third of the step — at 8192 envs the kernel is bound by the
// length of one thread's dependent instruction chain, not by bandwidth, so the chain is cut in three.
//   role 0 (base):   counters, derived base state, push / external-force windows, termination, base terms
//   role 1 (joints): the per-DOF reductions and the joint / distance terms, DOF-lag push
//   role 2 (feet):   feet Euler angles, contact bookkeeping (air time, clearance) and the feet terms
// The unscaled terms meet in shared memory; role 0 then forms the reward sum in the reference's
// (alphabetical) order and all threads share the reset bookkeeping.
constexpr int POST_ROLES = 3;

// named barrier between the substep workers (arrive) and role 1 (sync) of the fused kernel
constexpr int BAR_TORQUES = 1;
__device__ __forceinline__ void bar_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
// a store the compiler may not drop although the same address is stored again later in the kernel: every substep's
// torques land in memory in turn, as they do when the substeps are separate launches (lr:401-403)
__device__ __forceinline__ void store4_kept(float* ptr, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// lr:393-434 for one (env, group of four DOFs) when no simulator runs between the substeps: the action clip, then DEC x
// [_compute_torques, DOF-lag push] with the joint state, gains and offsets read once.  Same arithmetic, same Philox
// counters and the same stores as DEC launches of substep_kernel (ti5_substep.cu); the IMU-lag pushes, whose values equal
// the derived base state, are left to role 0.  DEC > 0: unrolled, all lagged rows in flight together; DEC == 0: any
// decimation, one row at a time.  Leaves the clipped actions and the last substep's torques in the CTA's tile.
template <int DEC>
__device__ __forceinline__ void substep_worker(const Ti5Params& p, const Ti5Buffers& b, const Ti5Rng& r,
                                               const float* __restrict__ actions_in, int64_t step, int idx, int e, int gq,
                                               float* tile_act, float* tile_tau) {
  const int N = p.num_envs, d0 = 4 * gq;
  const int dec = DEC > 0 ? DEC : p.decimation;
  const bool rg = p.flags & TI5_F_RAND_GAINS, fric = p.flags & TI5_F_RAND_COULOMB, rt = p.flags & TI5_F_RAND_TORQUE;
  const bool lagged = p.flags & TI5_F_ADD_LAG, philox = p.rng_mode == TI5_RNG_PHILOX;
  auto ld4 = [&](const float* base_ptr) { return reinterpret_cast<const float4*>(base_ptr)[idx]; };
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // ---- loads, all independent ----------------------------------------------------------------------
  const float4* ds = reinterpret_cast<const float4*>(b.dof_state);
  const float4 s0 = ds[2 * idx], s1 = ds[2 * idx + 1];
  float4 a4 = ld4(actions_in);
  const float4 off4 = ld4(b.motor_offsets);
  float4 kp4 = zero4, kd4 = zero4, vis4 = zero4, cou4 = zero4;
  int lag = 0;
  int64_t stamp = 0;
  if (lagged) { lag = b.lag_timestep[e * 3 + 0]; stamp = b.ring_stamp[e]; }
  if (rg) { kp4 = ld4(b.p_gains_r); kd4 = ld4(b.d_gains_r); }
  if (fric) { vis4 = ld4(b.viscous); cou4 = ld4(b.coulomb); }
  const int64_t base = (step - 1) * p.decimation;          // pushes completed before this step
  float4* ring = reinterpret_cast<float4*>(b.act_ring);
  const size_t ring_row = (size_t)N * 3;
  // ring slots of this step's pushes: one 64-bit remainder per ring, then 32-bit wrap-arounds
  const int as0 = (int)(base % p.lag_len), ds0 = (int)(base % p.dof_lag_len);
  auto wrap = [](int s, int len) { return s % len; };
  // the lagged action rows of the substeps that look back past the start of this step (lr:1045); rows pushed before the
  // env's last reset read as zero (lr:606).  Read before any row of this step is pushed: a slot this step overwrites
  // is only ever read by an EARLIER substep than the one that overwrites it (slot(base + k') == slot(base + k - lag)
  // needs k' = k + len - lag > k).
  float4 old_row[DEC > 0 ? DEC : 1];
  if (DEC > 0) {
#pragma unroll
    for (int k = 0; k < DEC; ++k) {
      const int64_t jj = base + k - lag;
      old_row[k] = zero4;
      if (lagged && lag > k && jj >= stamp && jj >= 0) old_row[k] = ring[(size_t)ring_slot(jj, p.lag_len) * ring_row + idx];
    }
  }
  a4 = make_float4(clampf(a4.x, -p.clip_actions, p.clip_actions), clampf(a4.y, -p.clip_actions, p.clip_actions),
                   clampf(a4.z, -p.clip_actions, p.clip_actions), clampf(a4.w, -p.clip_actions, p.clip_actions));   // lr:393-394
  reinterpret_cast<float4*>(b.actions)[idx] = a4;
  const float q[4] = {s0.x, s0.z, s1.x, s1.z}, qd[4] = {s0.y, s0.w, s1.y, s1.w};
  const float a[4] = {a4.x * p.action_scale, a4.y * p.action_scale, a4.z * p.action_scale, a4.w * p.action_scale};
  const float4 as4 = make_float4(a[0], a[1], a[2], a[3]);
  float kp[4] = {kp4.x, kp4.y, kp4.z, kp4.w}, kd[4] = {kd4.x, kd4.y, kd4.z, kd4.w};
  if (!rg) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { kp[i] = p.p_gains[d0 + i]; kd[i] = p.d_gains[d0 + i]; }
  }
  const float off[4] = {off4.x, off4.y, off4.z, off4.w};
  const float vis[4] = {vis4.x, vis4.y, vis4.z, vis4.w}, cou[4] = {cou4.x, cou4.y, cou4.z, cou4.w};
  const float4 q4 = make_float4(q[0], q[1], q[2], q[3]), qd4 = make_float4(qd[0], qd[1], qd[2], qd[3]);
  float tau[4] = {0.f, 0.f, 0.f, 0.f};
  auto substep = [&](int k, float4 t4) {
    // lr:1019-1074 torque of substep k
    float4 u4 = zero4;
    if (rt) u4 = philox ? philox_u4(p.seed, (uint64_t)step, S_TORQUE + k, idx)
                        : reinterpret_cast<const float4*>(r.torque)[(size_t)k * ring_row + idx];
    if (lagged) ring[(size_t)wrap(as0 + k, p.lag_len) * ring_row + idx] = as4;
    const bool from_ring = lagged && lag > 0;
    const float target[4] = {from_ring ? t4.x : a[0], from_ring ? t4.y : a[1], from_ring ? t4.z : a[2], from_ring ? t4.w : a[3]};
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
      float m[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        m[i] = affine(p.torque_multi_w, p.torque_multi_lo, u[i]);
        tau[i] = tau[i] * m[i];
      }
      store4_kept(b.torque_multi + (size_t)idx * 4, make_float4(m[0], m[1], m[2], m[3]));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float lim = p.torque_limits[d0 + i];
      tau[i] = clampf(tau[i], -lim, lim);
    }
    store4_kept(b.torques + (size_t)idx * 4, make_float4(tau[0], tau[1], tau[2], tau[3]));
    // lr:412-418 DOF-lag push after (what would be) simulator substep k
    if (p.flags & TI5_F_ADD_DOF_LAG) {
      float* row = b.dof_ring + ((size_t)wrap(ds0 + k, p.dof_lag_len) * N + e) * (2 * D);
      *reinterpret_cast<float4*>(row + d0) = q4;
      *reinterpret_cast<float4*>(row + D + d0) = qd4;
    }
  };
  if (DEC > 0) {
#pragma unroll
    for (int k = 0; k < DEC; ++k) substep(k, lag <= k ? as4 : old_row[k]);       // rows this step pushed: from registers
  } else {
#pragma unroll 1
    for (int k = 0; k < dec; ++k) {
      const int64_t jj = base + k - lag;
      float4 t4 = lag <= k ? as4 : zero4;
      if (lagged && lag > k && jj >= stamp && jj >= 0) t4 = ring[(size_t)ring_slot(jj, p.lag_len) * ring_row + idx];
      substep(k, t4);
    }
  }
  // what role 1 needs of this phase
  *reinterpret_cast<float4*>(tile_act + d0) = a4;
  *reinterpret_cast<float4*>(tile_tau + d0) = make_float4(tau[0], tau[1], tau[2], tau[3]);
}

// FUSED: the kernel also carries the CTA's envs through the DEC substeps that precede the post-physics phase
// (ti5_fused_step): 3 x TB further threads, one per (env, four DOFs), run them while the roles work; they meet role 1
// on a named barrier in front of the two terms over this step's actions and torques.
template <bool FUSED>
__global__ void __launch_bounds__(POST_ROLES * 128)
post_physics_kernel(const __grid_constant__ Ti5Params p, const __grid_constant__ Ti5Buffers b,
                    const __grid_constant__ Ti5Rng r, const __grid_constant__ PostSrc src,
                    const float* __restrict__ actions_in, int options) {
  chain_trigger();                                        // ti5_reset_observe may become resident
  const int push_last = options & TI5_POST_PUSH_LAST;
  const int N = p.num_envs;
  const int TB = p.env_block, tid = threadIdx.x;
  const int role = tid / TB, le = tid - role * TB;
  const int e0 = blockIdx.x * TB, e = e0 + le, n_tile = min(TB, N - e0);
  const bool live = e < N && role < POST_ROLES;
  Ti5Globals* g = b.globals;
  const int64_t step = g->step_index + 1;                 // index of the step in progress
  const int64_t counter = step + g->common_step_offset;   // common_step_counter after lr:471
  const bool first_force = g->is_first_add_force[step & 1] != 0;
  const bool philox = p.rng_mode == TI5_RNG_PHILOX;
  const int dm = p.div_mode;
  const uint32_t mask = p.term_mask;

  // window predicates are uniform over the grid (t1:193-215)
  bool push_window = false, force_window = false;
  if (p.flags & TI5_F_PUSH_ROBOTS) {
    int64_t i = counter / p.push_update_step;
    if (i >= p.n_push_dur) i = p.n_push_dur - 1;
    push_window = (double)(counter % p.push_interval) <= p.push_duration[i];
  }
  if (p.flags & TI5_F_ADD_EXT_FORCE) {
    int64_t i = counter / p.add_update_step;
    if (i >= p.n_add_dur) i = p.n_add_dur - 1;
    force_window = (double)(counter % p.ext_force_interval) <= p.add_duration[i];
  }

  probe(b.debug_ts, 0, 0);
  // ---- stage the CTA's tile of inputs: one TMA bulk copy per array, all in flight together ----------
  extern __shared__ __align__(128) unsigned char post_smem[];
  PostTile T;
  T.base = post_smem;
  T.tb = TB;
  T.src = &src;
  T.sums = reinterpret_cast<float*>(post_smem + (size_t)src.off[POST_CHUNKS] * TB);
  T.vals = T.sums + TI5_NUM_TERMS * TB;
  T.bar = reinterpret_cast<uint64_t*>(T.vals + TI5_NUM_TERMS * TB);
  // Early mode (chained launch, full tile): of everything this kernel reads, only `actions` and `torques` come from the
  // substep kernels of this step — every other array is simulator state (untouched while no simulator runs between
  // the kernels) or was last written by the previous step.  Roles 0 and 2 and most of role 1 therefore run BEFORE the
  // grid wait, on the first mbarrier; role 1 then waits for the substeps, pulls the two late arrays on the second
  // mbarrier and finishes the two terms that need them.  Nothing stored in front of the wait is read or written by a
  // substep kernel (root_states is: with push_robots the kernel keeps the plain order).
  // Small grids only (<= 12288 envs): there the CTAs are resident long before the substeps finish (common
  // carve-out, ti5_host.h) and the work in front of the wait is free; on large grids the plain order measured no worse.
  const bool early = !FUSED && (options & TI5_POST_CHAINED) && n_tile == TB && N <= TI5_SMALL_GRID_ENVS && !(p.flags & TI5_F_PUSH_ROBOTS);
  // FUSED: the two arrays come from this CTA's own substep workers (below), never from memory
  const bool late_by_tma = !FUSED && !early;
  const uint32_t late_bytes = (uint32_t)TB * (uint32_t)(src.rowb[C_ACT] + src.rowb[C_TORQUES]);
  auto issue_late = [&]() {      // one thread: arm the second barrier and start the two copies
    mbar_expect_tx(T.bar + 1, late_bytes);
    tma_load_1d(T.base + (size_t)src.off[C_ACT] * TB, static_cast<const char*>(src.ptr[C_ACT]) + (size_t)e0 * src.rowb[C_ACT],
                (uint32_t)(TB * src.rowb[C_ACT]), T.bar + 1);
    tma_load_1d(T.base + (size_t)src.off[C_TORQUES] * TB, static_cast<const char*>(src.ptr[C_TORQUES]) + (size_t)e0 * src.rowb[C_TORQUES],
                (uint32_t)(TB * src.rowb[C_TORQUES]), T.bar + 1);
  };
  if (n_tile == TB) {
    if (tid == 0) { mbar_init(T.bar, 1); mbar_init(T.bar + 1, 1); }
    __syncthreads();
    // thread k issues bulk copy k of the table; threads 32..59 one episode-sum column each; thread 0 arms the
    // barrier with the byte total (arrival order between the copies and the arm does not matter)
    // (the (TERMS, N) episode-sum columns start 16-byte aligned only when N is a multiple of 4)
    const bool sums_by_tma = (N & 3) == 0;
    if (tid < POST_CHUNKS) {
      if (tid != C_ACT && tid != C_TORQUES)
        tma_load_1d(T.base + (size_t)src.off[tid] * TB, static_cast<const char*>(src.ptr[tid]) + (size_t)e0 * src.rowb[tid],
                    (uint32_t)(TB * src.rowb[tid]), T.bar);
    } else if (sums_by_tma && tid >= 32 && tid < 32 + TI5_NUM_TERMS && (mask & (1u << (tid - 32)))) {
      const int t = tid - 32;
      tma_load_1d(T.sums + (size_t)t * TB, b.episode_sums + (size_t)t * N + e0, (uint32_t)(TB * 4), T.bar);
    }
    if (tid == 0)
      mbar_expect_tx(T.bar, (uint32_t)TB * (uint32_t)(src.off[POST_CHUNKS] + (sums_by_tma ? 4 * __popc(mask) : 0)) - late_bytes);
    if (!sums_by_tma) {
#pragma unroll 1
      for (int t = 0; t < TI5_NUM_TERMS; ++t)
        if (mask & (1u << t)) coop_load(T.sums + (size_t)t * TB, b.episode_sums + (size_t)t * N + e0, (uint32_t)(TB * 4));
    }
    if (late_by_tma) {
      chain_wait();                                       // the substep kernels are done
      if (tid == TB) issue_late();
    }
  } else {      // partial last tile: byte counts need not be multiples of 16, copy word by word
    if (!FUSED) chain_wait();
#pragma unroll 1
    for (int k = 0; k < POST_CHUNKS; ++k)
      if (!FUSED || (k != C_ACT && k != C_TORQUES))
        coop_load(T.base + (size_t)src.off[k] * TB, static_cast<const char*>(src.ptr[k]) + (size_t)e0 * src.rowb[k],
                  (uint32_t)(n_tile * src.rowb[k]));
#pragma unroll 1
    for (int t = 0; t < TI5_NUM_TERMS; ++t)
      if (mask & (1u << t)) coop_load(T.sums + (size_t)t * TB, b.episode_sums + (size_t)t * N + e0, (uint32_t)(n_tile * 4));
  }
  __syncthreads();
  if (FUSED) {
    if (role >= POST_ROLES) {
      // ======== substep workers: one thread per (env, four DOFs), coalesced over the CTA's 3 x TB groups ========
      const int item = tid - POST_ROLES * TB, wl = item / 3, gq = item - wl * 3;
      if (e0 + wl < N) {
        float* ta = T.at<float>(C_ACT) + wl * D;
        float* tt = T.at<float>(C_TORQUES) + wl * D;
        if (p.decimation == 10) substep_worker<10>(p, b, r, actions_in, step, e0 * 3 + item, e0 + wl, gq, ta, tt);
        else substep_worker<0>(p, b, r, actions_in, step, e0 * 3 + item, e0 + wl, gq, ta, tt);
      }
      bar_arrive(BAR_TORQUES, (POST_ROLES + 1) * TB);   // role 1 may read the two tile rows (stores above are ordered by it)
    }
    // a preceding ti5_sample_heights (chained launch) must have completed before this grid does: ti5_reset_observe,
    // which reads the heights, only waits for THIS grid
    if (options & TI5_FUSED_CHAINED) chain_wait();
  }
  if (n_tile == TB && role < POST_ROLES) {
    mbar_wait(T.bar, 0);
    if (late_by_tma) mbar_wait(T.bar + 1, 0);
  }
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

  if (live) {
    // ======== common prologue (every role, same values): command schedule, gait phase, contacts ==========
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
      cmd.z = mz ? affine((float)(cr[2][1] - cr[2][0]), (float)cr[2][0], u[2]) : 0.0f;
    }
    const float cmd_norm = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y + cmd.z * cmd.z);
    const bool stand = cmd_norm <= p.stand_threshold;
    // gait phase and stance mask (t1:80-107).  Side effect: standing envs restart the phase.
    if (stand) phase_len = 0;
    const float phase = (py_mod1(sdiv((float)phase_len * p.dt, p.cycle_time, dm)) + t_gait_start[le]) * (stand ? 0.0f : 1.0f);
    const float sin_pos = sinf(TWO_PI_F * phase);
    float stance[2] = {sin_pos >= 0.0f ? 1.0f : 0.0f, sin_pos < 0.0f ? 1.0f : 0.0f};
    if (fabsf(sin_pos) < 0.1f) stance[0] = stance[1] = 1.0f;
    const float* cf0 = t_contact + ((size_t)le * NB + p.feet[0]) * 3;
    const float* cf1 = t_contact + ((size_t)le * NB + p.feet[1]) * 3;
    const bool contact[2] = {cf0[2] > 5.0f, cf1[2] > 5.0f};
    const float* qrow = t_dof + (size_t)le * 2 * D;                           // interleaved (q, qd)

    if (role == 0) {
      // ================================ role 0: the base ===============================================
      float root[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) root[i] = t_root[le * RB + i];
      const float* tf = t_contact + ((size_t)le * NB + p.term_body) * 3;
      const float term_force = sqrtf(tf[0] * tf[0] + tf[1] * tf[1] + tf[2] * tf[2]);
      const float* pf = t_contact + ((size_t)le * NB + p.pen_body) * 3;
      const float pen_force = sqrtf(pf[0] * pf[0] + pf[1] * pf[1] + pf[2] * pf[2]);
      b.episode_length_buf[e] = ep_len;
      b.phase_length_buf[e] = phase_len;
      reinterpret_cast<float4*>(b.commands)[e] = cmd;
      // lr:475-479 derived base state
      const float bq[4] = {root[3], root[4], root[5], root[6]};
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
        const int s0 = (int)(((step - 1) * p.decimation) % p.imu_lag_len);
#pragma unroll 1
        for (int k = 0; k < p.decimation; ++k) {
          float* row = b.imu_ring + ((size_t)((s0 + k) % p.imu_lag_len) * N + e) * 6;
          row[0] = ang.x; row[1] = ang.y; row[2] = ang.z; row[3] = eul[0]; row[4] = eul[1]; row[5] = eul[2];
        }
      } else if (push_last && (p.flags & TI5_F_ADD_IMU_LAG)) {               // fused IMU-lag push of the last substep
        const int64_t j = (step - 1) * p.decimation + (p.decimation - 1);
        float* row = b.imu_ring + ((size_t)ring_slot(j, p.imu_lag_len) * N + e) * 6;
        row[0] = ang.x; row[1] = ang.y; row[2] = ang.z; row[3] = eul[0]; row[4] = eul[1]; row[5] = eul[2];
      }
      // t1:193-203, 217-231 push window: overwrite the base velocity
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
      // ---- base terms
      if (mask & (1u << T_BASE_ACC)) {                      // t1:717-724
        float sq = 0.0f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const float d = t_last_root_vel[le * 6 + i] - root[7 + i];
          sq += d * d;
        }
        put(T_BASE_ACC, expf_call(-sqrtf(sq) * 3.0f));
      }
      if (mask & (1u << T_COLLISION)) put(T_COLLISION, 1.0f * (pen_force > 0.1f ? 1.0f : 0.0f));   // t1:870-875
      if (mask & (1u << T_LOW_SPEED)) {                     // t1:816-847 (appendix A13)
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
        const float a = expf_call(-(fabsf(eul[0]) + fabsf(eul[1])) * 10.0f);
        const float bb = expf_call(-sqrtf(grav.x * grav.x + grav.y * grav.y) * 20.0f);
        put(T_ORIENTATION, (a + bb) / 2.0f);
      }
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
    } else if (role == 1) {
      // ================================ role 1: the joints =============================================
      if (!FUSED && push_last && (p.flags & TI5_F_ADD_DOF_LAG)) {            // fused DOF-lag push of the last substep
        const int64_t j = (step - 1) * p.decimation + (p.decimation - 1);
        float* row = b.dof_ring + ((size_t)ring_slot(j, p.dof_lag_len) * N + e) * (2 * D);
#pragma unroll 1
        for (int i = 0; i < D; ++i) { row[i] = qrow[2 * i]; row[D + i] = qrow[2 * i + 1]; }
      }
      const float* ldv = t_last_dof_vel + le * D;
      const float* ref = t_ref + le * D;
      // one pass over the 12 DOFs feeds every per-DOF reduction that needs joint state only (each sum keeps its own
      // DOF order); the sums over `actions` and `torques` follow below, behind the grid wait in early mode
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
      const float* f0 = t_rigid + ((size_t)le * NB + p.feet[0]) * RB;
      const float* f1 = t_rigid + ((size_t)le * NB + p.feet[1]) * RB;
      const float* k0 = t_rigid + ((size_t)le * NB + p.knees[0]) * RB;
      const float* k1 = t_rigid + ((size_t)le * NB + p.knees[1]) * RB;
      if (mask & (1u << T_FEET_DISTANCE))                   // t1:599-612
        put(T_FEET_DISTANCE, pair_distance_reward(f0[0], f0[1], f1[0], f1[1], p.foot_min_dist, p.foot_max_dist));
      if (mask & (1u << T_KNEE_DISTANCE))                   // t1:615-628
        put(T_KNEE_DISTANCE, pair_distance_reward(k0[0], k0[1], k1[0], k1[1], p.knee_min_dist, p.knee_max_dist));
    } else {
      // ================================ role 2: the feet ===============================================
      FootState foot[2];
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        const float* rs = t_rigid + ((size_t)le * NB + p.feet[f]) * RB;
        foot[f].pos[0] = rs[0]; foot[f].pos[1] = rs[1]; foot[f].pos[2] = rs[2];
        foot[f].quat[0] = rs[3]; foot[f].quat[1] = rs[4]; foot[f].quat[2] = rs[5]; foot[f].quat[3] = rs[6];
        foot[f].wxy[0] = rs[10]; foot[f].wxy[1] = rs[11];
        const float* cf = f == 0 ? cf0 : cf1;
        foot[f].force[0] = cf[0]; foot[f].force[1] = cf[1]; foot[f].force[2] = cf[2];
        float fe[3];
        euler_xyz(foot[f].quat, fe);                                           // lr:480-481
        foot[f].pitch = fe[1];
#pragma unroll
        for (int i = 0; i < 3; ++i) b.feet_euler_xyz[e * 6 + 3 * f + i] = fe[i];
      }
      if (mask & (1u << T_BASE_HEIGHT)) {                   // t1:706-715
        const float ground = (foot[0].pos[2] * stance[0] + foot[1].pos[2] * stance[1]) / (stance[0] + stance[1]);
        const float h = t_root[le * RB + 2] - (ground - 0.05f);
        put(T_BASE_HEIGHT, expf_call(-fabsf(h - p.base_height_target) * 100.0f));
      }
      if (mask & (1u << T_FEET_AIR_TIME)) {                 // t1:642-657 (appendix A6, A8)
        const bool tiny = cmd_norm < 0.05f;
        float air_sum = 0.0f;
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const float st = tiny ? 1.0f : stance[f];
          const bool filt = contact[f] || (st != 0.0f) || (t_last_contacts[le * 2 + f] != 0);
          b.contact_filt[e * 2 + f] = filt ? 1 : 0;
          b.last_contacts[e * 2 + f] = contact[f] ? 1 : 0;
          float air = t_air[le * 2 + f];
          const float first = (air > 0.0f && filt) ? 1.0f : 0.0f;
          air += p.dt;
          air_sum += clampf(air, 0.0f, 0.5f) * first;
          b.feet_air_time[e * 2 + f] = air * (filt ? 0.0f : 1.0f);
        }
        put(T_FEET_AIR_TIME, air_sum);
      }
      if (mask & (1u << T_FEET_CLEARANCE)) {                // t1:793-814 (appendix A9)
        float sw = 0.0f;
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const float z = foot[f].pos[2];
          const float h = t_feet_h[le * 2 + f] + (z - t_last_z[le * 2 + f]);
          b.last_feet_z[e * 2 + f] = z;
          const float swing = 1.0f - stance[f];
          const float hit = (h > p.target_feet_height && h < p.target_feet_height_max) ? 1.0f : 0.0f;
          sw += hit * swing;
          b.feet_height[e * 2 + f] = h * (contact[f] ? 0.0f : 1.0f);
        }
        put(T_FEET_CLEARANCE, sw);
      }
      if (mask & (1u << T_FEET_CONTACT_FORCES)) {           // t1:679-684
        float sq = 0.0f;
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const float n = sqrtf((foot[f].force[0] * foot[f].force[0] + foot[f].force[1] * foot[f].force[1]) +
                                foot[f].force[2] * foot[f].force[2]);
          sq += clampf(n - p.max_contact_force, 0.0f, 400.0f);
        }
        put(T_FEET_CONTACT_FORCES, sq);
      }
      if (mask & (1u << T_FEET_CONTACT_NUMBER)) {           // t1:659-668 (appendix A14)
        float sq = 0.0f;
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const float st = stand ? 1.0f : stance[f];
          sq += ((contact[f] ? 1.0f : 0.0f) == st) ? 1.0f : -0.3f;
        }
        put(T_FEET_CONTACT_NUMBER, sq / 2.0f);
      }
      if (mask & (1u << T_FEET_ROTATION)) {                 // t1:926-935 (appendix A11)
        const float rot = foot[0].pitch * foot[0].pitch + foot[1].pitch * foot[1].pitch;
        const float x = rot / 1.0f;
        put(T_FEET_ROTATION, 1.0f * expf_call(-(x * x)));
      }
      if (mask & (1u << T_FEET_STUMBLE)) {                  // t1:937-940
        bool any = false;
#pragma unroll
        for (int f = 0; f < 2; ++f)
          any = any || (sqrtf(foot[f].force[0] * foot[f].force[0] + foot[f].force[1] * foot[f].force[1]) >
                        5.0f * fabsf(foot[f].force[2]));
        put(T_FEET_STUMBLE, any ? 1.0f : 0.0f);
      }
      if (mask & (1u << T_FOOT_SLIP)) {                     // t1:630-640 (appendix A10)
        float sq = 0.0f;
#pragma unroll
        for (int f = 0; f < 2; ++f)
          sq += sqrtf(sqrtf(foot[f].wxy[0] * foot[f].wxy[0] + foot[f].wxy[1] * foot[f].wxy[1])) * (contact[f] ? 1.0f : 0.0f);
        put(T_FOOT_SLIP, sq);
      }
      if (mask & (1u << T_STAND_STILL)) {                   // t1:899-915 (appendix A12)
        // dof_idx [0,1,2,3,5,6,7,8] with weights [2,2,1,1,1,2,2,1], then the two foot pitches with weight 1
        float sq = 0.0f;
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          const int j = i < 4 ? i : i + 1;
          const float w = (i < 2 || i == 5 || i == 6) ? 2.0f : 1.0f;
          const float er = (qrow[2 * j] - p.default_dof_pos[j]) * w;
          sq += er * er;
        }
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const float er = foot[f].pitch * 1.0f;
          sq += er * er;
        }
        put(T_STAND_STILL, stand ? expf_call(-sq) : 0.0f);
      }
    }
  }
  // ---- role 1, last: the two terms over this step's actions and torques ---------------------------------------
  if (role == 1) {
    if (FUSED) bar_sync(BAR_TORQUES, (POST_ROLES + 1) * TB);     // the CTA's substep workers have left both rows in the tile
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
  probe(b.debug_ts, 0, 6, TB);          // role 1 done
  probe(b.debug_ts, 0, 7, 2 * TB);      // role 2 done
  __syncthreads();
  probe(b.debug_ts, 0, 3);

  // ---- lr:654-680: reward sum in alphabetical term order, per-term episode sums, clip at zero -----------
  // The per-term episode sums are independent of one another: roles 1 and 2 take every other term while role 0
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
  reset_bookkeeping(p, b, reset, e, le, T.sums, TB, step);
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
  reset_bookkeeping(p, b, reset, e, threadIdx.x, nullptr, 0, step);
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
  TI5_CHECK_ARGS(p->rng_mode == TI5_RNG_PHILOX || (r && r->cmd));
  TI5_CHECK_ARGS((p->term_mask & (1u << T_DOF_VEL_LIMITS)) == 0);   // the reference term reads a cfg field t1 lacks
  TI5_CHECK_ARGS(!(p->flags & TI5_F_ADD_EXT_FORCE) || p->applied_stride >= 3);
  Ti5Rng rr = r ? *r : Ti5Rng{};
  const int blocks = (p->num_envs + p->env_block - 1) / p->env_block;
  const PostSrc src = make_post_src(*p, *b);
  const size_t smem = post_tile_bytes(p->env_block, src.off[POST_CHUNKS]);
  auto kernel = fused ? post_physics_kernel<true> : post_physics_kernel<false>;
  if (!ti5_ensure_smem(kernel, smem)) {
    ti5_set_error("%s: %zu bytes of shared memory per CTA not available", what, smem);
    return TI5_ECUDA;
  }
  ti5_set_carveout(kernel, ti5_small_grid(p));
  const int threads = (fused ? 2 : 1) * POST_ROLES * p->env_block;
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
