// Library-level entry points: version, error reporting, struct sizes, term names, and the
// stand-alone reset-index compaction (lr:490 `reset_buf.nonzero(as_tuple=False).flatten()`).
#include <cstdarg>
#include <cstdio>

#include "ti5_device.cuh"
#include "ti5_host.h"

static thread_local char g_err[512] = "";

void ti5_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ti5_check_launch(const char* what) {
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    ti5_set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(err));
    return TI5_ECUDA;
  }
  return TI5_OK;
}

extern "C" int ti5_version(void) { return TI5_ABI_VERSION; }
extern "C" const char* ti5_last_error(void) { return g_err; }

extern "C" int ti5_struct_sizes(int32_t out[4]) {
  if (!out) return TI5_EINVAL;
  out[0] = (int32_t)sizeof(Ti5Params);
  out[1] = (int32_t)sizeof(Ti5Buffers);
  out[2] = (int32_t)sizeof(Ti5Rng);
  out[3] = (int32_t)sizeof(Ti5Globals);
  return TI5_OK;
}

extern "C" int ti5_rollout_struct_sizes(int32_t out[3]) {
  if (!out) return TI5_EINVAL;
  out[0] = (int32_t)sizeof(Ti5Rollout);
  out[1] = (int32_t)sizeof(Ti5Transition);
  out[2] = (int32_t)sizeof(Ti5Batch);
  return TI5_OK;
}

static const char* kTermNames[TI5_NUM_TERMS] = {
    "action_smoothness", "base_acc", "base_height", "collision", "default_joint_pos", "dof_acc", "dof_vel",
    "dof_vel_limits", "feet_air_time", "feet_clearance", "feet_contact_forces", "feet_contact_number",
    "feet_distance", "feet_rotation", "feet_stumble", "foot_slip", "joint_pos", "knee_distance", "low_speed",
    "orientation", "stand_still", "stand_sysmetry", "termination", "torques", "track_vel_hard",
    "tracking_ang_vel", "tracking_lin_vel", "vel_mismatch_exp"};

extern "C" const char* ti5_reward_name(int term) {
  return (term >= 0 && term < TI5_NUM_TERMS) ? kTermNames[term] : nullptr;
}

namespace ti5 {

constexpr int CB = 1024;   // mask bytes per CTA of the stand-alone compaction

// pass 1: per-CTA popcount; the last CTA to finish turns the counts into exclusive offsets
__global__ void __launch_bounds__(CB) compact_count_kernel(const uint8_t* __restrict__ mask, int n, int* scratch,
                                                           int* count_out) {
  __shared__ int s_warp[32];
  __shared__ bool s_last;
  const int i = blockIdx.x * CB + threadIdx.x;
  const bool f = i < n && mask[i] != 0;
  const BlockRank br = block_rank(f, s_warp);
  int* counts = scratch + 1;            // scratch[0] is the ticket
  if (threadIdx.x == 0) {
    counts[blockIdx.x] = br.total;
    __threadfence();
    s_last = atomicAdd(scratch, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    int run = 0;
    for (int blk = 0; blk < (int)gridDim.x; ++blk) {
      const int c = ((volatile int*)counts)[blk];
      counts[blk] = run;
      run += c;
    }
    *count_out = run;
    scratch[0] = 0;
  }
}

// pass 2: ascending scatter
__global__ void __launch_bounds__(CB) compact_scatter_kernel(const uint8_t* __restrict__ mask, int n,
                                                             const int* __restrict__ scratch, int* __restrict__ ids) {
  __shared__ int s_warp[32];
  const int i = blockIdx.x * CB + threadIdx.x;
  const bool f = i < n && mask[i] != 0;
  const BlockRank br = block_rank(f, s_warp);
  if (f) ids[scratch[1 + blockIdx.x] + br.rank] = i;
}

// lr:915-939: per-env joint properties of the listed envs, dense, for one hand-over to the simulator
__global__ void __launch_bounds__(256) gather_dof_props_kernel(const float* __restrict__ armatures, const int32_t* __restrict__ ids,
                                                               const int32_t* __restrict__ count, int capacity, int flags,
                                                               float* __restrict__ out, const float* __restrict__ coeffs,
                                                               int flags2) {
  const int n = min(*count, capacity);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * D; i += gridDim.x * blockDim.x) {
    const int r = i / D, d = i - r * D;
    float* o = out + (size_t)i * 3;
    // friction / damping multiplier of the env (randomize_joint_friction / _damping: off in t1_cfg), lr:921-930
    o[0] = (flags2 & TI5_F2_RAND_JOINT_FRICTION) ? coeffs[(size_t)ids[r] * 2 + 0] : 1.0f;
    o[1] = (flags2 & TI5_F2_RAND_JOINT_DAMPING) ? coeffs[(size_t)ids[r] * 2 + 1] : 1.0f;
    o[2] = (flags & TI5_F_RAND_ARMATURE) ? armatures[(size_t)ids[r] * D + d] : 0.0f;
  }
}

}  // namespace ti5

extern "C" int ti5_gather_dof_props(const Ti5Params* p, const Ti5Buffers* b, const int32_t* ids, const int32_t* count,
                                    int32_t capacity, float* props_out, void* stream) {
  TI5_CHECK_ARGS(p && b && ids && count && props_out && capacity > 0 && b->joint_armatures);
  TI5_CHECK_ARGS(p->flags2 == 0 || b->joint_coeffs);
  const int blocks = min((capacity * ti5::D + 255) / 256, ti5_sm_count() * 4);
  ti5::gather_dof_props_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(b->joint_armatures, ids, count, capacity, p->flags, props_out,
                                                                         b->joint_coeffs, p->flags2);
  return ti5_check_launch("ti5_gather_dof_props");
}

extern "C" int ti5_compact_resets(const uint8_t* mask, int32_t n, int32_t* ids_out, int32_t* count_out,
                                  int32_t* scratch, void* stream) {
  TI5_CHECK_ARGS(count_out && n >= 0);
  if (n == 0) {   // empty mask: nothing to read, zero ids

    cudaMemsetAsync(count_out, 0, sizeof(int32_t), (cudaStream_t)stream);
    return ti5_check_launch("ti5_compact_resets");
  }
  TI5_CHECK_ARGS(mask && ids_out && scratch);
  const int blocks = (n + ti5::CB - 1) / ti5::CB;
  ti5::compact_count_kernel<<<blocks, ti5::CB, 0, (cudaStream_t)stream>>>(mask, n, scratch, count_out);
  ti5::compact_scatter_kernel<<<blocks, ti5::CB, 0, (cudaStream_t)stream>>>(mask, n, scratch, ids_out);
  return ti5_check_launch("ti5_compact_resets");
}
