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

// Optional common shared-memory carve-out (TI5_CARVEOUT = percent) for the kernels of a chained step.  An SM only
// hosts CTAs of kernels that agree on its shared memory / L1 split; with a common carve-out the CTAs of the next
// kernels become resident several launches ahead (measured: post_physics CTAs resident 12 us before their inputs
// exist) — but the step is bound by the dependency chain, not by CTA launch, and it measured 0.3 us slower.  Off by
// default; kept as a knob for other grid sizes.
#include <cstdlib>
template <class K>
inline void ti5_set_carveout(K kernel) {
  static const int pct = getenv("TI5_CARVEOUT") ? atoi(getenv("TI5_CARVEOUT")) : -1;
  if (pct >= 0) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}
