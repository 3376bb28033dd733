// Terrain height sampling around every robot: lr:1551-1587 `_get_heights` with
// humanoid/utils/math.py:8-12 `quat_apply_yaw` and isaacgym `normalize` / `quat_apply`.
// One warp per env, lanes over its scan points: the (N, 187) output is written coalesced; the three int16
// gathers per point hit the 8.8 MB height field, which stays L2-resident.
#include "ti5_device.cuh"
#include "ti5_host.h"

namespace ti5 {

// One warp per env: the yaw quaternion, its normalisation (a square root and two IEEE divisions) and the base position
// are per-env quantities — computed once per warp instead of once per scan point (ncu, round 2: 220 instructions per
// point, 75 % issue-active, i.e. bound by exactly that redundant arithmetic) — then the lanes stride over the env's 187
// points: coalesced (N,187) output, three int16 gathers per point from the 8.8 MB (L2-resident) height field.
constexpr int HEIGHT_WARPS = 8;      // envs per CTA
__global__ void __launch_bounds__(HEIGHT_WARPS * 32)
heights_kernel(const __grid_constant__ Ti5Params p, const __grid_constant__ Ti5Buffers b) {
  chain_trigger();                       // the step kernel that follows needs nothing from this grid before its end
  const int npts = p.num_height_points;
  const int e = blockIdx.x * HEIGHT_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= p.num_envs) return;
  float* out = b.measured_heights + (size_t)e * npts;
  if (p.flags & TI5_F_PLANE) {           // lr:1564-1565
    for (int k = lane; k < npts; k += 32) out[k] = 0.0f;
    return;
  }
  const float* root = b.root_states + (size_t)e * RB;
  // yaw-only quaternion, normalised: x / norm.clamp(min=1e-9)
  const float qz = root[5], qw = root[6], rx = root[0], ry = root[1];
  float n = sqrtf(((0.0f * 0.0f + 0.0f * 0.0f) + qz * qz) + qw * qw);
  n = n < 1e-9f ? 1e-9f : n;
  const float z = qz / n, w = qw / n;
  const int16_t* hs = b.height_samples;
  const int rows = p.height_rows, cols = p.height_cols;
  // Per point the kernel is bound by instruction issue (ncu at 65536 envs: 78 % issue-active, 88 instructions per point),
  // so the loop body carries only the operations whose rounding reaches the cell index:
  //  * quat_apply(q, v) = v + w * t + cross(q.xyz, t), t = 2 * cross(q.xyz, v) with q.xyz = (0, 0, z), v.z = 0: the
  //    products with the zero components are +-0 for the finite grid offsets and are left out — `x + (+-0)` and
  //    `(+-0) - x` round like `x` and `-x`; only the sign of an exactly-zero sum could differ, and a zero of either sign
  //    truncates to cell 0;
  //  * `.long()` (truncation toward zero) followed by the clip to [0, dim - 2]: the quotient is clamped to [-1, dim - 1]
  //    as a float first (exact: dim < 2^24), so a 32-bit conversion truncates it like the 64-bit one; below -1 / above
  //    dim - 1 / NaN end on the same side of the clip as before (NaN: fmaxf returns -1 -> cell 0, as cvt of NaN did).
  const float xhi = (float)(rows - 1), yhi = (float)(cols - 1);
  const float scale = p.horizontal_scale, inv_scale = 1.0f / scale;      // ATen's CUDA `x / scalar` multiplies by this
  const bool recip = p.div_mode == TI5_DIV_RECIPROCAL;
  const float2* pts = reinterpret_cast<const float2*>(b.height_points);
#pragma unroll 2
  for (int k = lane; k < npts; k += 32) {
    const float2 v = pts[k];
    const float vx = v.x, vy = v.y;
    const float tx = (0.0f - z * vy) * 2.0f;
    const float ty = (z * vx) * 2.0f;
    float px = (vx + w * tx) + (0.0f - z * ty);
    float py = (vy + w * ty) + (z * tx);
    px = (px + rx) + p.border_size;
    py = (py + ry) + p.border_size;
    const float qx = recip ? px * inv_scale : px / scale, qy = recip ? py * inv_scale : py / scale;
    int ix = (int)fminf(fmaxf(qx, -1.0f), xhi);               // .long() truncates toward zero
    int iy = (int)fminf(fmaxf(qy, -1.0f), yhi);
    ix = max(0, min(ix, rows - 2));
    iy = max(0, min(iy, cols - 2));
    const int o = ix * cols + iy;
    const int16_t h1 = hs[o];
    const int16_t h2 = hs[o + cols];
    const int16_t h3 = hs[o + 1];
    int16_t h = h1 < h2 ? h1 : h2;
    h = h < h3 ? h : h3;
    out[k] = (float)h * p.vertical_scale;
  }
}

}  // namespace ti5

extern "C" int ti5_sample_heights(const Ti5Params* p, const Ti5Buffers* b, void* stream) {
  TI5_CHECK_ARGS(p && b && p->num_envs > 0 && p->num_height_points > 0 && b->measured_heights);
  TI5_CHECK_ARGS((p->flags & TI5_F_PLANE) || (b->height_samples && b->height_points && p->height_rows >= 2 && p->height_cols >= 2));
  TI5_CHECK_ARGS((int64_t)p->height_rows * p->height_cols < (1ll << 31) && p->height_rows < (1 << 24) && p->height_cols < (1 << 24));
  TI5_CHECK_ARGS((p->flags & TI5_F_PLANE) || ((uintptr_t)b->height_points & 7) == 0);      // read as (x, y) pairs
  ti5::heights_kernel<<<(p->num_envs + ti5::HEIGHT_WARPS - 1) / ti5::HEIGHT_WARPS, ti5::HEIGHT_WARPS * 32, 0, (cudaStream_t)stream>>>(*p, *b);
  return ti5_check_launch("ti5_sample_heights");
}
