// Micro-benchmark: per-node time of back-to-back tiny kernels inside a CUDA graph on this GPU
// (the floor any per-substep launch pays), for small and 2 KB kernel-parameter blocks.
#include <cstdio>
#include <cuda_runtime.h>
struct Big { char bytes[2048]; };
__global__ void k_small(float* p) { if (threadIdx.x == 0 && blockIdx.x == 0 && p) p[0] += 1.f; }
__global__ void k_big(const __grid_constant__ Big b, float* p) { if (threadIdx.x == 0 && blockIdx.x == 0 && p) p[0] += b.bytes[5]; }
__global__ void k_touch(const __grid_constant__ Big b, float* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = p[i] * 1.0001f + b.bytes[3];
}
template <class F> float run(F launch, int nodes, int reps) {
  cudaStream_t s; cudaStreamCreate(&s);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < nodes; ++i) launch(s);
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 5; ++i) cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  cudaEventRecord(a, s);
  for (int i = 0; i < reps; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(b, s); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms * 1e3f / (reps * nodes);
}
int main() {
  float* d; cudaMalloc(&d, 98304 * 4); cudaMemset(d, 0, 98304 * 4);
  Big big = {};
  printf("empty kernel, 8 B params        : %.2f us/node\n", run([&](cudaStream_t s) { k_small<<<1, 32, 0, s>>>(d); }, 11, 200));
  printf("empty kernel, 2 KB params       : %.2f us/node\n", run([&](cudaStream_t s) { k_big<<<1, 32, 0, s>>>(big, d); }, 11, 200));
  printf("384x64 empty, 2 KB params       : %.2f us/node\n", run([&](cudaStream_t s) { k_big<<<384, 64, 0, s>>>(big, d); }, 11, 200));
  printf("384x64 RMW of 96 K floats, 2 KB : %.2f us/node\n", run([&](cudaStream_t s) { k_touch<<<384, 64, 0, s>>>(big, d, 98304); }, 11, 200));
  printf("96x256 RMW of 96 K floats, 2 KB : %.2f us/node\n", run([&](cudaStream_t s) { k_touch<<<96, 256, 0, s>>>(big, d, 98304); }, 11, 200));
  return 0;
}
