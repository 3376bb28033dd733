// Device-side helpers shared by the ti5 step kernels (sm_100a).
//
// Arithmetic follows the reference's torch fp32 op chains one rounding at a time: the
// library is built with -fmad=false so that a*b+c rounds twice like two eager torch ops,
// libdevice sinf/cosf/expf/atan2f/asinf/sqrtf are the functions torch's CUDA kernels call,
// and `tensor / python_scalar` is rounded the way the reference's device does (div_mode).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ti5_step.h"

namespace ti5 {

constexpr int D = TI5_NUM_DOF;
constexpr int NB = TI5_NUM_BODIES;
constexpr int RB = 13;  // floats per rigid-body / root state row

// `2 * torch.pi` and `np.pi` as torch rounds them when they meet an fp32 tensor
constexpr float TWO_PI_F = 6.283185307179586f;
constexpr float PI_F = 3.141592653589793f;
constexpr float HALF_PI_F = 1.5707963267948966f;

// reward term indices: alphabetical order of every `_reward_*` of the task (t1:576-946),
// which is the order `dir()` yields and the reference sums in (appendix A1)
enum Term {
  T_ACTION_SMOOTHNESS = 0, T_BASE_ACC, T_BASE_HEIGHT, T_COLLISION, T_DEFAULT_JOINT_POS, T_DOF_ACC, T_DOF_VEL,
  T_DOF_VEL_LIMITS, T_FEET_AIR_TIME, T_FEET_CLEARANCE, T_FEET_CONTACT_FORCES, T_FEET_CONTACT_NUMBER,
  T_FEET_DISTANCE, T_FEET_ROTATION, T_FEET_STUMBLE, T_FOOT_SLIP, T_JOINT_POS, T_KNEE_DISTANCE, T_LOW_SPEED,
  T_ORIENTATION, T_STAND_STILL, T_STAND_SYSMETRY, T_TERMINATION, T_TORQUES, T_TRACK_VEL_HARD,
  T_TRACKING_ANG_VEL, T_TRACKING_LIN_VEL, T_VEL_MISMATCH_EXP
};
static_assert(T_VEL_MISMATCH_EXP == TI5_NUM_TERMS - 1, "term table out of sync with TI5_NUM_TERMS");

// extras_log columns after the TI5_NUM_TERMS episode means
constexpr int LOG_TERRAIN_LEVEL = TI5_NUM_TERMS;
constexpr int LOG_MAX_COMMAND_X = TI5_NUM_TERMS + 1;
constexpr int LOG_N_RESET = TI5_NUM_TERMS + 2;

// ---------------------------------------------------------------------------------------------
// scalar helpers
// ---------------------------------------------------------------------------------------------

// tensor / python scalar (see TI5_DIV_*)
__device__ __forceinline__ float sdiv(float x, float c, int mode) {
  return mode == TI5_DIV_RECIPROCAL ? x * (1.0f / c) : x / c;
}

// torch.remainder for floats (python-style modulo)
__device__ __forceinline__ float py_mod(float a, float b) {
  float m = fmodf(a, b);
  if (m != 0.0f && ((b < 0.0f) != (m < 0.0f))) m += b;
  return m;
}

// a % 1.0 (the gait phase, t1:88): fmodf(a, 1) == a - trunc(a) exactly in fp32 (Sterbenz for |a| >= 1, trunc = 0 below;
// integers beyond 2^23 give 0, inf gives NaN either way) — two instructions instead of libdevice's reduction loop
__device__ __forceinline__ float py_mod1(float a) {
  float m = a - truncf(a);
  if (m != 0.0f && m < 0.0f) m += 1.0f;
  return m;
}

// torch.clip / clamp: min(max(x, lo), hi) with NaN passed through
__device__ __forceinline__ float clampf(float x, float lo, float hi) {
  return x < lo ? lo : (x > hi ? hi : x);
}

__device__ __forceinline__ float signf(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }

// torch_rand_float(lo, hi): (hi - lo) * u + lo with the width rounded once (host) to fp32
__device__ __forceinline__ float affine(float w, float lo, float u) { return w * u + lo; }

struct V3 {
  float x, y, z;
};

// isaacgym.torch_utils.quat_rotate_inverse, q = (x, y, z, w)
static __device__ __noinline__ V3 quat_rotate_inverse(const float q[4], V3 v) {
  const float qw = q[3];
  const float s = 2.0f * (qw * qw) - 1.0f;
  const float cx = q[1] * v.z - q[2] * v.y;
  const float cy = q[2] * v.x - q[0] * v.z;
  const float cz = q[0] * v.y - q[1] * v.x;
  const float dot = q[0] * v.x + q[1] * v.y + q[2] * v.z;
  V3 o;
  o.x = (v.x * s - cx * qw * 2.0f) + q[0] * dot * 2.0f;
  o.y = (v.y * s - cy * qw * 2.0f) + q[1] * dot * 2.0f;
  o.z = (v.z * s - cz * qw * 2.0f) + q[2] * dot * 2.0f;
  return o;
}

// lr:27-53 get_euler_xyz_tensor: one wrapped angle (which = 0 roll, 1 pitch, 2 yaw)
// The argument is an atan2f / asinf result, |a| <= pi < 2 pi, where fmodf(a, 2 pi) returns a itself: the python-style
// remainder reduces to "add 2 pi to negatives" (-0.0 and NaN pass through like in fmodf), then the reference's fold
// back — (a + 2 pi) - 2 pi in fp32, not a, bit for bit as torch computes it.
__device__ __forceinline__ float wrap_angle(float a) {
  if (a < 0.0f) a += TWO_PI_F;
  if (a > PI_F) a -= TWO_PI_F;
  return a;
}
__device__ __forceinline__ float euler_roll(const float q[4]) {
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  return wrap_angle(atan2f(2.0f * (w * x + y * z), ((w * w - x * x) - y * y) + z * z));
}
__device__ __forceinline__ float euler_pitch(const float q[4]) {
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  const float sinp = 2.0f * (w * y - z * x);
  const float pitch = fabsf(sinp) >= 1.0f ? fabsf(HALF_PI_F) * signf(sinp) : asinf(sinp);
  return wrap_angle(pitch);
}
__device__ __forceinline__ float euler_yaw(const float q[4]) {
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  return wrap_angle(atan2f(2.0f * (w * z + x * y), ((w * w + x * x) - y * y) - z * z));
}
// t1:185-188 (heading mode): forward = quat_apply(base_quat, (1, 0, 0)) as isaacgym.torch_utils writes it
// (t = 2 * cross(q.xyz, v); v + w * t + cross(q.xyz, t), every product formed: signed zeros come out as in torch),
// heading = atan2(forward.y, forward.x), yaw rate = clip(0.5 * wrap_to_pi(target - heading), -1, 1) with
// wrap_to_pi of humanoid/utils/math.py:15-18: torch's `%` (fmod, then + b when the signs differ), then
// `-= float32(2 pi) * (angle > float32(pi))`.  Out of line: off in t1_cfg, never on the hot path's instruction stream.
#ifdef TI5_NO_HEADING                  // A/B builds only: what the (uniform, never taken in t1_cfg) heading branches cost
#define TI5_HEADING(p) false
#else
#define TI5_HEADING(p) (((p).flags & TI5_F_HEADING_COMMAND) != 0)
#endif
static __device__ __noinline__ float heading_yaw_rate(const float q[4], float target) {
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  const float vx = 1.0f, vy = 0.0f, vz = 0.0f;
  const float tx = (y * vz - z * vy) * 2.0f, ty = (z * vx - x * vz) * 2.0f, tz = (x * vy - y * vx) * 2.0f;
  const float fx = (vx + w * tx) + (y * tz - z * ty);
  const float fy = (vy + w * ty) + (z * tx - x * tz);
  const float heading = atan2f(fy, fx);
  float m = fmodf(target - heading, TWO_PI_F);
  if (m != 0.0f && m < 0.0f) m += TWO_PI_F;
  m -= TWO_PI_F * (m > PI_F ? 1.0f : 0.0f);
  return clampf(0.5f * m, -1.0f, 1.0f);
}

// expf as a real call: ~25 call sites share one copy of the libdevice body (and its I-cache lines)
static __device__ __noinline__ float expf_call(float x) { return expf(x); }

static __device__ __noinline__ void euler_xyz(const float q[4], float e[3]) {
  e[0] = euler_roll(q);
  e[1] = euler_pitch(q);
  e[2] = euler_yaw(q);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (throughput mode).  One call yields four 32-bit words for the counter
// (idx, site, step) under the key `seed`; words -> U[0,1) with 24 random bits.
// ---------------------------------------------------------------------------------------------
enum RngSite { S_TORQUE = 0, S_CMD = 16, S_PUSH = 24, S_EXT, S_DOFS, S_ROOT, S_DR, S_GAIT_TIME, S_NOISE, S_LAG,
               S_GAIT_START, S_TERRAIN,
               S_LAGSTEP };   // per-step lag re-draws and the position / velocity lag draws (idx = env * 32 + slot: 0-15 the
                              // action lag of substep k, 17-20 DOF / IMU / position / velocity per step, 24-25 reset draws)
// the options t1_cfg marks "always False": tested together, so the common path pays one uniform branch
#define TI5_F_LAG_OPTIONS (TI5_F_LAG_PERSTEP | TI5_F_DOF_LAG_PERSTEP | TI5_F_IMU_LAG_PERSTEP | TI5_F_POS_VEL_LAG | \
                           TI5_F_POS_LAG_PERSTEP | TI5_F_VEL_LAG_PERSTEP)

__device__ __forceinline__ uint4 philox4_inline(uint64_t seed, uint64_t step, uint32_t site, uint32_t idx) {
  uint32_t c0 = idx, c1 = site, c2 = (uint32_t)step, c3 = (uint32_t)(step >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// out-of-line copy for the rarely taken draw sites (commands, pushes, resets): one body, many callers
static __device__ __noinline__ uint4 philox4(uint64_t seed, uint64_t step, uint32_t site, uint32_t idx) {
  return philox4_inline(seed, step, site, idx);
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }
__device__ __forceinline__ float philox_u(uint64_t seed, uint64_t step, uint32_t site, uint32_t idx) {
  const uint4 r = philox4(seed, step, site, idx >> 2);
  const uint32_t lane = idx & 3u;
  return u01(lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w);
}

// four uniforms of one Philox call: element group `idx4` of a site
__device__ __forceinline__ float4 philox_u4(uint64_t seed, uint64_t step, uint32_t site, uint32_t idx4) {
  const uint4 r = philox4(seed, step, site, idx4);
  return make_float4(u01(r.x), u01(r.y), u01(r.z), u01(r.w));
}

// ---------------------------------------------------------------------------------------------
// vector loads of per-env rows
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load12(const float* __restrict__ base, int e, float v[12]) {
  const float4* p = reinterpret_cast<const float4*>(base + (size_t)e * 12);
  const float4 a = p[0], b = p[1], c = p[2];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
}
__device__ __forceinline__ void store12(float* __restrict__ base, int e, const float v[12]) {
  float4* p = reinterpret_cast<float4*>(base + (size_t)e * 12);
  p[0] = make_float4(v[0], v[1], v[2], v[3]);
  p[1] = make_float4(v[4], v[5], v[6], v[7]);
  p[2] = make_float4(v[8], v[9], v[10], v[11]);
}

// ---------------------------------------------------------------------------------------------
// Tile staging: global -> shared with the TMA engine (cp.async.bulk + mbarrier).  A CTA that owns a
// contiguous block of envs pulls every per-env array it needs as one bulk copy per array: all of the
// tile's bytes are in flight at once and the math then runs out of shared memory, instead of ~100
// scattered, serially exposed global loads per thread.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned) completing on `bar`
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// one 4-byte asynchronous copy global -> shared (LDGSTS): for rows the bulk engine cannot take (52-byte rows of an
// AoS tensor are neither 16-byte aligned nor a multiple of 16 long); the issuing thread does not wait for the data
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// one 16-byte asynchronous copy global -> shared (both addresses 16-byte aligned), L2 only
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
// fallback for a partial tile (byte count not a multiple of 16): the CTA's threads copy word by word
__device__ __forceinline__ void coop_load(void* dst, const void* src, uint32_t bytes) {
  const uint32_t words = bytes >> 2;
  for (uint32_t i = threadIdx.x; i < words; i += blockDim.x)
    reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  const uint32_t tail = bytes & 3u;     // byte arrays (bool masks)
  if (threadIdx.x < tail)
    reinterpret_cast<uint8_t*>(dst)[(words << 2) + threadIdx.x] = reinterpret_cast<const uint8_t*>(src)[(words << 2) + threadIdx.x];
}

// the same by a subset of the CTA: threads 0 .. nthreads-1, `t` = the caller's index among them
__device__ __forceinline__ void coop_load_n(void* dst, const void* src, uint32_t bytes, int t, int nthreads) {
  const uint32_t words = bytes >> 2;
  for (uint32_t i = t; i < words; i += nthreads)
    reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  const uint32_t tail = bytes & 3u;
  if ((uint32_t)t < tail)
    reinterpret_cast<uint8_t*>(dst)[(words << 2) + t] = reinterpret_cast<const uint8_t*>(src)[(words << 2) + t];
}

static __device__ __noinline__ void coop_load_call(void* dst, const void* src, uint32_t bytes) { coop_load(dst, src, bytes); }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch ("chained" launches of the fused step): a kernel launched with the
// programmatic-serialization attribute may become resident while its predecessor on the stream is still
// running.  chain_trigger() lets the successor's CTAs be scheduled; chain_wait() blocks until the predecessor
// grid has completed and its writes are visible.  Everything a kernel does before chain_wait() must therefore
// be independent of every other kernel of the step (loads of simulator tensors and of state that only the
// previous STEP wrote, address arithmetic, random draws); every store comes after it.  Both are no-ops for a
// launch without the attribute.
// ---------------------------------------------------------------------------------------------
// L2 prefetch: legal in front of chain_wait() for ANY address (it returns nothing), useful for rows that were written
// in earlier steps and have been evicted since
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void chain_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void chain_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// profiling aid: nanosecond timestamp probe `k` of this CTA (thread 0 only), when a probe buffer is bound
// (compiled in with -DTI5_PROBES only — tools/probe.py builds its own library: ten predicated probe sites were 2.4 % of
// the fused kernel's executed instructions)
#ifdef TI5_PROBES
__device__ __forceinline__ void probe(uint64_t* ts, int kernel, int k, int thread = 0) {
  if (ts != nullptr && (int)threadIdx.x == thread) {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    ts[((size_t)kernel * 4096 + blockIdx.x) * 8 + k] = t;
  }
}
#else
__device__ __forceinline__ void probe(uint64_t*, int, int, int = 0) {}
#endif

// 64-bit step counters divided by run-time divisors: a 64-bit division is a ~100-instruction subroutine with a long
// dependent chain, and the kernels are bound by exactly that (one warp runs the whole program once).  Counters fit
// 31 bits for any realistic run (2^31 steps = 8 months at 100 steps/s): one 32-bit division then; the 64-bit path stays
// for correctness beyond.
__device__ __forceinline__ int64_t fast_mod(int64_t x, int64_t m) {
  if ((((uint64_t)x | (uint64_t)m) >> 31) == 0) return (int64_t)((uint32_t)x % (uint32_t)m);
  return x % m;
}
__device__ __forceinline__ int64_t fast_div(int64_t x, int64_t m) {
  if ((((uint64_t)x | (uint64_t)m) >> 31) == 0) return (int64_t)((uint32_t)x / (uint32_t)m);
  return x / m;
}
// ring slot of push index j (j >= 0)
__device__ __forceinline__ int ring_slot(int64_t j, int len) { return (int)fast_mod(j, len); }

// block-wide exclusive scan helper result for compaction
struct BlockRank {
  int rank;   // exclusive rank of this thread's flag within the CTA
  int total;  // flags set in the CTA
};
__device__ __forceinline__ BlockRank block_rank(bool flag, int* s_warp /* >= 32 ints */) {
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  int before = 0, total = 0;
  for (int w = 0; w < nwarp; ++w) {
    const int c = s_warp[w];
    if (w < warp) before += c;
    total += c;
  }
  __syncthreads();
  BlockRank r;
  r.rank = before + __popc(bal & ((1u << lane) - 1u));
  r.total = total;
  return r;
}

}  // namespace ti5
