// Terrain height sampling around every robot: lr:1551-1587 `_get_heights` with
// humanoid/utils/math.py:8-12 `quat_apply_yaw` and isaacgym `normalize` / `quat_apply`.
// One thread per (env, scan point): the (N, 187) output is written coalesced; the three int16
// gathers per point hit the 8.8 MB height field, which stays L2-resident.
#include "ti5_device.cuh"
#include "ti5_host.h"

namespace ti5 {

__global__ void __launch_bounds__(256)
heights_kernel(const __grid_constant__ Ti5Params p, const __grid_constant__ Ti5Buffers b) {
  chain_trigger();                       // the step kernel that follows needs nothing from this grid before its end
  const int npts = p.num_height_points;
  const size_t total = (size_t)p.num_envs * npts;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (p.flags & TI5_F_PLANE) {           // lr:1564-1565
    b.measured_heights[i] = 0.0f;
    return;
  }
  const int e = (int)(i / npts), k = (int)(i - (size_t)e * npts);
  const float* root = b.root_states + (size_t)e * RB;
  // yaw-only quaternion, normalised: x / norm.clamp(min=1e-9)
  const float qz = root[5], qw = root[6];
  float n = sqrtf(((0.0f * 0.0f + 0.0f * 0.0f) + qz * qz) + qw * qw);
  n = n < 1e-9f ? 1e-9f : n;
  const float z = qz / n, w = qw / n;
  const float vx = b.height_points[k * 2 + 0], vy = b.height_points[k * 2 + 1];
  // quat_apply(q, v) = v + w * t + cross(q.xyz, t),  t = 2 * cross(q.xyz, v),  q.xyz = (0, 0, z), v.z = 0
  const float tx = (0.0f * 0.0f - z * vy) * 2.0f;
  const float ty = (z * vx - 0.0f * 0.0f) * 2.0f;
  const float tz = (0.0f * vy - 0.0f * vx) * 2.0f;
  float px = (vx + w * tx) + (0.0f * tz - z * ty);
  float py = (vy + w * ty) + (z * tx - 0.0f * tz);
  px = (px + root[0]) + p.border_size;
  py = (py + root[1]) + p.border_size;
  long long ix = (long long)sdiv(px, p.horizontal_scale, p.div_mode);   // .long() truncates toward zero
  long long iy = (long long)sdiv(py, p.horizontal_scale, p.div_mode);
  ix = ix < 0 ? 0 : (ix > p.height_rows - 2 ? p.height_rows - 2 : ix);
  iy = iy < 0 ? 0 : (iy > p.height_cols - 2 ? p.height_cols - 2 : iy);
  const int16_t* hs = b.height_samples;
  const int16_t h1 = hs[ix * p.height_cols + iy];
  const int16_t h2 = hs[(ix + 1) * p.height_cols + iy];
  const int16_t h3 = hs[ix * p.height_cols + iy + 1];
  int16_t h = h1 < h2 ? h1 : h2;
  h = h < h3 ? h : h3;
  b.measured_heights[i] = (float)h * p.vertical_scale;
}

}  // namespace ti5

extern "C" int ti5_sample_heights(const Ti5Params* p, const Ti5Buffers* b, void* stream) {
  TI5_CHECK_ARGS(p && b && p->num_envs > 0 && p->num_height_points > 0 && b->measured_heights);
  TI5_CHECK_ARGS((p->flags & TI5_F_PLANE) || (b->height_samples && b->height_points && p->height_rows >= 2 && p->height_cols >= 2));
  const size_t total = (size_t)p->num_envs * p->num_height_points;
  ti5::heights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*p, *b);
  return ti5_check_launch("ti5_sample_heights");
}
