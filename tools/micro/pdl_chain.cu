// Micro-benchmark: ten dependent "substep-like" kernels (384 CTAs x 64 threads, one float4 read-modify-write per thread
// plus a dependent ALU chain) launched back to back:
//   plain      ordinary launches (grid-to-grid dependency)
//   plain+pub  the same, each CTA also publishing a progress flag (fence + release) at its end
//   pdl+wait   programmatic dependent launch, griddepcontrol.wait before the dependent load
//   pdl+flags  programmatic dependent launch, NO grid wait: each CTA polls the flag of the same CTA of its predecessor
// each in a CUDA graph and directly on a stream.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct Big { char bytes[1800]; };
enum { PUBLISH = 1, GRIDWAIT = 2, FLAGWAIT = 4, TRIGGER = 8 };
__global__ void __launch_bounds__(64) stage(const __grid_constant__ Big big, float4* data, long long* flags, long long epoch, int mode, int n) {
  if (mode & TRIGGER) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (mode & GRIDWAIT) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (mode & FLAGWAIT) {
    if (threadIdx.x == 0) {
      long long seen; unsigned spins = 0;
      do { asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(seen) : "l"(flags + blockIdx.x) : "memory");
           if (++spins > (1u << 26)) __trap(); } while (seen < epoch - 1);
    }
    __syncthreads();
  }
  float4 v = idx < n ? data[idx] : make_float4(0, 0, 0, 0);
  float x = v.x + big.bytes[7];
#pragma unroll 1
  for (int i = 0; i < 64; ++i) x = x * 1.000001f + 0.5f;      // ~64 dependent FMAs ~ 0.15 us
  v.x = x; v.y += 1.f;
  if (idx < n) data[idx] = v;
  if (mode & PUBLISH) {
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(flags + blockIdx.x), "l"(epoch) : "memory"); }
  }
}
static long long g_epoch = 0;
static void launch(cudaStream_t s, const Big& big, float4* d, long long* f, int mode, bool pdl, int n) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((n + 63) / 64); cfg.blockDim = dim3(64); cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, stage, big, d, f, 0LL, mode, n);
}
// epochs: flags are reset to 0 before each chain by a memset node; stage k publishes k+1 and waits for k
static void chain(cudaStream_t s, const Big& big, float4* d, long long* f, int mode, bool pdl, int n, int stages, int blocks) {
  if (mode & (PUBLISH | FLAGWAIT)) cudaMemsetAsync(f, 0, blocks * 8, s);
  for (int k = 0; k < stages; ++k) {
    const bool first = k == 0;
    int m = mode;
    if (first) m &= ~(GRIDWAIT | FLAGWAIT);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(64); cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (pdl && !first) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, stage, big, d, f, (long long)(k + 1), m, n);
  }
}
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 24576, stages = 10, blocks = (n + 63) / 64, reps = 300;
  float4* d; long long* f; cudaMalloc(&d, n * 16); cudaMalloc(&f, blocks * 8);
  Big big = {};
  struct { const char* name; int mode; bool pdl; } V[] = {
    {"plain", 0, false}, {"plain+pub", PUBLISH, false}, {"pdl+wait", GRIDWAIT | TRIGGER, true},
    {"pdl+flags", PUBLISH | FLAGWAIT | TRIGGER, true}, {"flags only (no pdl attr)", PUBLISH | FLAGWAIT, false}};
  cudaStream_t s; cudaStreamCreate(&s);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (auto& v : V) {
    for (int use_graph = 1; use_graph >= 0; --use_graph) {
      cudaMemset(d, 0, n * 16);
      cudaGraph_t g; cudaGraphExec_t ge = nullptr;
      if (use_graph) {
        cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
        chain(s, big, d, f, v.mode, v.pdl, n, stages, blocks);
        cudaStreamEndCapture(s, &g);
        if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) { printf("%s: instantiate failed: %s\n", v.name, cudaGetErrorString(cudaGetLastError())); continue; }
      }
      auto go = [&]() { if (use_graph) cudaGraphLaunch(ge, s); else chain(s, big, d, f, v.mode, v.pdl, n, stages, blocks); };
      for (int i = 0; i < 5; ++i) go();
      cudaStreamSynchronize(s);
      cudaEventRecord(a, s);
      for (int i = 0; i < reps; ++i) go();
      cudaEventRecord(b, s); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      float4 h; cudaMemcpy(&h, d + (n - 1), 16, cudaMemcpyDeviceToHost);
      cudaError_t err = cudaGetLastError();
      printf("%-26s %-6s : %6.2f us per 10-stage chain, %5.2f us/stage   check y=%.0f (want %d) %s\n", v.name, use_graph ? "graph" : "stream",
             ms * 1e3f / reps, ms * 1e3f / reps / stages, h.y, (reps + 5) * stages, err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
  }
  return 0;
}
