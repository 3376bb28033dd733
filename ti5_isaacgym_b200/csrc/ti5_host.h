// Host-side plumbing shared by the C-ABI translation units: argument checks, launch checks,
// the thread-local last-error string.
#pragma once
#include <cuda_runtime.h>

#include "../../include/ti5_step.h"

#include <utility>

void ti5_set_error(const char* fmt, ...);
int ti5_check_launch(const char* what);

#define TI5_CHECK_ARGS(cond)                                                        \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      ti5_set_error("%s: invalid argument: !(%s)", __func__, #cond);                \
      return TI5_EINVAL;                                                            \
    }                                                                               \
  } while (0)


// Kernel launch with an optional programmatic dependency on the preceding kernel of the stream (see
// chain_trigger / chain_wait in ti5_device.cuh).  Works under stream capture: the edge becomes a programmatic
// dependency of the CUDA graph.
template <class... KArgs, class... Args>
inline cudaError_t ti5_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream, bool chained,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = chained ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Common shared-memory carve-out for the kernels of a chained step on SMALL grids (<= 12288 envs).
// An SM only hosts CTAs of kernels that agree on its shared memory / L1 split; with a common carve-out the CTAs of
// the following kernels become resident several launches ahead (measured: ti5_post_physics CTAs resident 12 us before
// their inputs exist), which is what lets ti5_post_physics do most of its work in front of the grid wait (50.8 vs
// 53.9 us/step at 8192 envs).  On large grids the SMs are full anyway and the smaller L1 costs the substep kernel
// 20 % (236 vs 212 us/step at 65536 envs), so the kernels keep the default there.  TI5_CARVEOUT=<percent|-1> overrides.
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#define TI5_SMALL_GRID_ENVS 12288    /* measured: early mode + carve-out 58.8 vs 60.2 us at 12288 envs, 67.2 vs 66.5 at 16384 */
inline bool ti5_small_grid(const Ti5Params* p) { return p->num_envs <= TI5_SMALL_GRID_ENVS; }

// Function attributes are per DEVICE: the caches below are keyed by (current device, kernel) and guarded by a mutex, so
// a process that drives several GPUs (or several host threads) configures every kernel on every device it launches on.
struct Ti5AttrCache {
  std::mutex mu;
  std::map<std::pair<int, const void*>, int> carveout;
  std::map<std::pair<int, const void*>, size_t> smem;
  std::map<int, int> sms;
  std::map<std::tuple<int, const void*, int, size_t>, int> occupancy;
};
inline Ti5AttrCache& ti5_attr_cache() {
  static Ti5AttrCache c;
  return c;
}
inline int ti5_current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
template <class K>
inline void ti5_set_carveout(K kernel, bool small_grid) {
  static const int forced = getenv("TI5_CARVEOUT") ? atoi(getenv("TI5_CARVEOUT")) : -2;
  const int want = forced != -2 ? forced : (small_grid ? 100 : -1);
  Ti5AttrCache& c = ti5_attr_cache();
  const auto key = std::make_pair(ti5_current_device(), reinterpret_cast<const void*>(kernel));
  std::lock_guard<std::mutex> lock(c.mu);
  auto it = c.carveout.find(key);
  if (it == c.carveout.end() || it->second != want) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, want);
    c.carveout[key] = want;
  }
}
// opt a kernel in to `bytes` of dynamic shared memory on the current device (no-op once done for at least that much)
template <class K>
inline bool ti5_ensure_smem(K kernel, size_t bytes) {
  // The 48 KB a kernel may use without opting in cover its STATIC shared memory too: a dynamic size just below 48 KB (the
  // fused step with decimation 4: 47.6 KB + 1.1 KB static) fails to launch with "invalid argument" unless the attribute
  // is raised — found by the `plane_params` parity scenario.  Opt in from 32 KB on (static usage here is ~1 KB).
  if (bytes <= 32 * 1024) return true;
  Ti5AttrCache& c = ti5_attr_cache();
  const auto key = std::make_pair(ti5_current_device(), reinterpret_cast<const void*>(kernel));
  std::lock_guard<std::mutex> lock(c.mu);
  auto it = c.smem.find(key);
  if (it != c.smem.end() && it->second >= bytes) return true;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  c.smem[key] = bytes;
  return true;
}
// SM count of the current device (148 on a B200)
// resident CTAs per SM of `kernel` launched with `threads` threads and `smem` bytes of dynamic shared memory (cached)
template <class K>
inline int ti5_ctas_per_sm(K kernel, int threads, size_t smem) {
  Ti5AttrCache& c = ti5_attr_cache();
  const auto key = std::make_tuple(ti5_current_device(), reinterpret_cast<const void*>(kernel), threads, smem);
  std::lock_guard<std::mutex> lock(c.mu);
  auto it = c.occupancy.find(key);
  if (it != c.occupancy.end()) return it->second;
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = 1;
  }
  c.occupancy[key] = n;
  return n;
}
inline int ti5_sm_count() {
  Ti5AttrCache& c = ti5_attr_cache();
  const int dev = ti5_current_device();
  std::lock_guard<std::mutex> lock(c.mu);
  auto it = c.sms.find(dev);
  if (it != c.sms.end()) return it->second;
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  c.sms[dev] = n;
  return n;
}
